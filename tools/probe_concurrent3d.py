"""GPU probe: do several 3-D near-far solves overlap on one GPU?  K handles of the same grid, one host thread each
(ctypes releases the GIL), every thread solves its share of the sources into device tables.
usage: python tools/probe_concurrent3d.py [nn] [nsrc]"""
import ctypes as C
import json
import sys
import threading
import time

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import numpy as np
import torch
import rt_loader

rt = rt_loader.load()
R = rt.R
C0 = (np.deg2rad(70.0), np.deg2rad(70.0), R - 2000.0)
C1 = (np.deg2rad(110.0), np.deg2rad(110.0), R)
nn = int(sys.argv[1]) if len(sys.argv) > 1 else 216
nsrc = int(sys.argv[2]) if len(sys.argv) > 2 else 8
prof = rt.velocity_profile()
itp = rt.LinearInterpolation(prof.r, prof.Vp)
grids = [rt.grid(C0, C1, (nn, nn, nn), neighbour_levels=1, coord_system="spherical") for _ in range(4)]
n = grids[0].n
X, Y, Z = grids[0].coordinates()
U = torch.from_numpy(rt.interpolate_velocity(np.minimum(np.sqrt(X * X + Y * Y + Z * Z), R), itp)).cuda()
del X, Y, Z
lat = int(np.ceil(np.sqrt(nsrc)))
srcs = np.array([1 + (nn * (2 * (q % lat) + 1)) // (2 * lat) + nn * ((nn * (2 * (q // lat) + 1)) // (2 * lat) + nn * (nn - 1))
                 for q in range(nsrc)], np.int64)
d = torch.empty((nsrc, n), dtype=torch.float64, device="cuda")
p = torch.empty((nsrc, n), dtype=torch.int32, device="cuda")
for g in grids:
    g._handle.set_option("schedule", 1)
ref = None
for K in (1, 2, 4, 1):
    def work(k):
        mine = np.ascontiguousarray(srcs[k::K])
        for j, s in enumerate(mine):
            st = rt.RtStats()
            q = k + j * K
            rt.api.check(rt.lib().rt_bfm_solve_dev(grids[k]._handle.h, U.data_ptr(), np.array([s], np.int64), 1, 64,
                                                   d[q].data_ptr(), p[q].data_ptr(), C.byref(st)))
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(k,)) for k in range(K)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
    if ref is None:
        ref = d.clone()
    print(json.dumps(dict(nn=nn, nsrc=nsrc, concurrent=K, ms_total=round(dt, 2), ms_per_source=round(dt / nsrc, 2),
                          same=bool(torch.equal(ref, d)))), flush=True)
