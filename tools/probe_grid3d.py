"""GPU probe: 3-D shell grid (bench.py's grid3d workloads), Jacobi against the two near-far variants.
usage: python tools/probe_grid3d.py "216;368" "jacobi;pull:8;pull:4:0.5;push:8"   (variant[:delta_factor[:early_advance]])"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import numpy as np
import torch
import rt_loader

rt = rt_loader.load()
R = rt.R
C0 = (np.deg2rad(70.0), np.deg2rad(70.0), R - 2000.0)
C1 = (np.deg2rad(110.0), np.deg2rad(110.0), R)
prof = rt.velocity_profile()
itp = rt.LinearInterpolation(prof.r, prof.Vp)
sizes = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "64").split(";")]
modes = (sys.argv[2] if len(sys.argv) > 2 else "jacobi;pull:8;push:8").split(";")
for nn in sizes:
    g = rt.grid(C0, C1, (nn, nn, nn), neighbour_levels=int(os.environ.get("RT_LEVELS", "1")), coord_system="spherical")
    h, n = g._handle, g.n
    X, Y, Z = g.coordinates()
    U_host = rt.interpolate_velocity(np.minimum(np.sqrt(X * X + Y * Y + Z * Z), R), itp)
    del X, Y, Z
    U = torch.from_numpy(U_host).cuda()
    src = np.array([1 + (nn // 2) + nn * ((nn // 2) + nn * (nn - 1))], np.int64)
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    p = torch.empty(n, dtype=torch.int32, device="cuda")
    ref = None
    h.set_option("profile_timers", int(os.environ.get("RT_TIMERS", "0")))
    h.set_option("check_every", int(os.environ.get("RT_CHECK_EVERY", "0")))
    h.set_option("use_graph", int(os.environ.get("RT_GRAPH", "1")))
    h.set_option("fuse_begin", int(os.environ.get("RT_FUSE", "0")))
    for mode in modes:
        name, _, df = mode.partition(":")
        df, _, ea = df.partition(":")
        h.set_option("early_advance", float(ea) if ea else -1.0)
        h.set_option("schedule", 0 if name == "jacobi" else 1)
        h.set_option("tile_pull", 0 if name == "push" else 1)
        h.set_option("delta", 0.0)
        h.set_option("delta_factor", float(df) if df else 0.0)
        ts = []
        for rep in range(4):
            st = rt.RtStats()
            torch.cuda.synchronize()
            t = time.perf_counter()
            rt.api.check(rt.lib().rt_bfm_solve_dev(h.h, U.data_ptr(), src, 1, 64, d.data_ptr(), p.data_ptr(), C.byref(st)))
            ts.append((time.perf_counter() - t) * 1e3)
        sd = st.as_dict()
        if ref is None:
            ref = d.clone()
        print(json.dumps(dict(nn=nn, n=n, mode=mode, ms=round(float(np.median(ts[1:])), 3), first_ms=round(ts[0], 2),
                              rounds=sd["sweeps"], pull_rounds=sd["relax_launches"], relaxed=sd["relaxed_edges"],
                              ratio=round(sd["relaxed_edges"] / sd["graph_edges"], 3), releases=sd["vertex_updates"],
                              kernel_ms=round(sd["kernel_ms"], 3), relax_ms=round(sd["relax_ms"], 3),
                              prev_ms=round(sd["prev_ms"], 3), screened=sd["screened_edges"], exact=sd["exact_edges"], launches=sd["total_launches"],
                              same_dist=bool(torch.equal(ref, d)))), flush=True)
    del g, h, U, d, p, ref
