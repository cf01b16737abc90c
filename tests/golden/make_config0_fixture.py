"""Pins BASELINE configs[0] (the README example: annulus 180x50 at 1 km spacing, AK135, surface source theta = 0,
30 receiver paths) at FULL size with the CPU oracle.  One oracle solve at this size is ~1.3e12 candidate evaluations
(tens of minutes on 8 cores), so the result is committed as a small fixture: sha256 of the travel-time and predecessor
tables, the travel times at the 30 README receivers, the 30 reconstructed paths, the sweep count.

Like the other goldens this freezes the ORACLE (the reference is pure Julia and cannot run here); what it adds is a
full-size bit-exact anchor for the CUDA path in both schedules.  Run from the repository root (once):
    python tests/golden/make_config0_fixture.py [spacing_km]
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
R = 6371.0

if __name__ == "__main__":
    spacing = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    d = np.load(os.path.join(ROOT, "raytracer.jl_b200", "data", "ak135_profile.npz"))
    kr = (d["depth_km"].max() - d["depth_km"])[::-1].copy()
    kv = d["vp"][::-1].copy()
    t0 = time.time()
    m = O.Annulus(180, 50, spacing)
    U = O.interp_velocity(kr, kv, m.r)
    src = O.closest_point(m.theta, m.r, 0.0, R)
    print("mesh", m.n, "nodes; source", src, "; build %.1f s" % (time.time() - t0), flush=True)
    t0 = time.time()
    dist, prev, st = O.bfm(m, U, src, nthreads=O.num_threads())
    print("solve %.1f s, %d sweeps, %d evaluations" % (time.time() - t0, st["sweeps"], st["relaxed_edges"]), flush=True)
    degs = np.concatenate([np.arange(10, 151, 10), 360 - np.arange(150, 9, -10)]).astype(np.float32)  # README.md:42-45
    th = np.deg2rad(degs).astype(np.float64)
    recv = np.array([O.closest_point(m.theta, m.r, float(t), R) for t in th], np.int64)
    paths = [O.reconstruct_path(prev, src, int(r)) for r in recv]
    off = np.concatenate([[0], np.cumsum([len(p) for p in paths])]).astype(np.int64)
    name = "config0_180_50_%gkm.npz" % spacing
    np.savez_compressed(os.path.join(HERE, name), n=m.n, nel=m.nel, halo_rows=m.halo_rows, source=src,
                        sweeps=st["sweeps"], relaxed_edges=st["relaxed_edges"], graph_edges=st["graph_edges"],
                        sha256_dist=hashlib.sha256(np.ascontiguousarray(dist).tobytes()).hexdigest(),
                        sha256_prev=hashlib.sha256(np.ascontiguousarray(prev, np.int64).tobytes()).hexdigest(),
                        sha256_U=hashlib.sha256(np.ascontiguousarray(U).tobytes()).hexdigest(),
                        sha256_topology=hashlib.sha256(m.e2n_idx.tobytes() + m.G_rowval.tobytes() + m.halo.tobytes()).hexdigest(),
                        receivers=recv, T_receivers=dist[recv - 1], path_off=off, path_idx=np.concatenate(paths),
                        dist_sample_idx=np.arange(0, m.n, 997, dtype=np.int64) + 1, dist_sample=dist[::997])
    print("written", name, "T(30,90,150 deg) =", dist[recv[2] - 1], dist[recv[8] - 1], dist[recv[14] - 1])
