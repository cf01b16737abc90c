"""Regenerates the committed golden vectors from the CPU oracle.

The reference is pure Julia with no tests/fixtures and Julia is not installed, so these vectors are NOT
outputs of the reference itself: they freeze the oracle's restatement (regression pin) so that both the oracle
and the CUDA path are compared against fixed bytes.  Run from the repository root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
R = 6371.0

if __name__ == "__main__":
    d = np.load(os.path.join(ROOT, "raytracer.jl_b200", "data", "ak135_profile.npz"))
    kr = (d["depth_km"].max() - d["depth_km"])[::-1].copy()
    kv = d["vp"][::-1].copy()
    m = O.Annulus(24, 6, 300.0)
    U = O.interp_velocity(kr, kv, m.r)
    src = O.closest_point(m.theta, m.r, 0.0, R)
    dist, prev, st = O.bfm(m, U, src)
    np.savez_compressed(os.path.join(HERE, "annulus_24_6_300.npz"), sizes=np.array([m.n, m.nel, m.halo_rows]),
                        e2n_off=m.e2n_off, e2n_idx=m.e2n_idx, G_colptr=m.G_colptr, G_rowval=m.G_rowval,
                        halo=m.halo, theta=m.theta, r=m.r, x=m.x, z=m.z, U=U, source=src, dist=dist, prev=prev,
                        sweeps=st["sweeps"])
    # the other relax modes on the same mesh: Float32 arithmetic (bfm_gpu.jl) and the dual-velocity relax (bfm.jl:113-159)
    d32, p32, s32 = O.bfm_f32(m, U, src)
    V2 = O.dual_velocity(kr, kv, m.r, 1.0)
    dd, pd, sd = O.bfm_dual(m, V2, src)
    np.savez_compressed(os.path.join(HERE, "annulus_24_6_300_modes.npz"), dist_f32=d32.astype(np.float32), prev_f32=p32,
                        sweeps_f32=s32["sweeps"], V2=V2, dist_dual=dd, prev_dual=pd, sweeps_dual=sd["sweeps"])
    c0 = np.array([np.deg2rad(70.0), np.deg2rad(70.0), R - 2000.0])
    c1 = np.array([np.deg2rad(110.0), np.deg2rad(110.0), R])
    X, Y, Z = O.grid3d_coords(c0, c1, (7, 6, 5), 1)
    rr = np.sqrt(X * X + Y * Y + Z * Z)
    U3 = O.interp_velocity(kr, kv, np.minimum(rr, R))
    d3, p3, s3 = O.bfm3d((7, 6, 5), 1, X, Y, Z, U3, 18)
    np.savez_compressed(os.path.join(HERE, "grid3d_7_6_5.npz"), c0=c0, c1=c1, X=X, Y=Y, Z=Z, U=U3, source=18,
                        dist=d3, prev=p3, sweeps=s3["sweeps"])
    d3f, p3f, _ = O.bfm3d_f32((7, 6, 5), 1, X, Y, Z, U3, 18)
    np.savez_compressed(os.path.join(HERE, "grid3d_7_6_5_f32.npz"), dist_f32=d3f.astype(np.float32), prev_f32=p3f)
    print("golden vectors written:", m.n, "nodes 2-D;", len(X), "nodes 3-D")
