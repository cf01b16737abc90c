"""GPU probe: cost of the 1000-receiver path sweep after a 3-D solve in either near-far variant."""
import ctypes as C
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
nn = int(sys.argv[1]) if len(sys.argv) > 1 else 216
g = rt.grid(C0, C1, (nn, nn, nn), neighbour_levels=1, coord_system="spherical")
h, n = g._handle, g.n
U = torch.full((n,), 8.0, dtype=torch.float64, device="cuda")
th = np.linspace(C0[0], C1[0], 40)
ph = np.linspace(C0[1], C1[1], 25)
TH, PH = np.meshgrid(th, ph, indexing="ij")
recv = np.asarray(rt.closest_point(g, TH.reshape(-1), PH.reshape(-1), np.full(TH.size, R)), np.int64)
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 1
src = np.array([1 + (nn // 2) + k + nn * ((nn // 2) + nn * (nn - 1)) for k in range(ns)], np.int64)
d = torch.empty((ns, n), dtype=torch.float64, device="cuda")
p = torch.empty((ns, n), dtype=torch.int32, device="cuda")
h.set_option("schedule", 1)
for tp in (0, 1, 1):
    h.set_option("tile_pull", tp)
    st = rt.RtStats()
    rt.api.check(rt.lib().rt_bfm_solve_dev(h.h, U.data_ptr(), src, ns, 64, d.data_ptr(), p.data_ptr(), C.byref(st)))
    torch.cuda.synchronize()
    ts = []
    for rep in range(3):
        t0 = time.perf_counter()
        off = np.zeros(len(recv) + 1, np.int64)
        rt.api.check(rt.lib().rt_reconstruct_paths_dev(p[ns - 1].data_ptr(), n, int(src[ns - 1]), recv, len(recv), off, None, 0))
        t1 = time.perf_counter()
        idx = np.zeros(int(off[-1]), np.int64)
        rt.api.check(rt.lib().rt_reconstruct_paths_dev(p[ns - 1].data_ptr(), n, int(src[ns - 1]), recv, len(recv), off, idx.ctypes.data, len(idx)))
        t2 = time.perf_counter()
        ts.append((round((t1 - t0) * 1e3, 2), round((t2 - t1) * 1e3, 2)))
    print("tile_pull", tp, "path nodes", int(off[-1]), "max len", int(np.diff(off).max()), "ms (len, fill) x3:", ts, flush=True)
