"""GPU probe: a batch of sources on the 3-D shell grid through ONE handle: sequential (batch = 1) against the
per-source slots (batch = 0 / k).   usage: python tools/probe_batch3d.py [nn] [nsrc] ["1;2;4;0"]"""
import ctypes as C
import json
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
nsrc = int(sys.argv[2]) if len(sys.argv) > 2 else 8
modes = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "1;2;4;0").split(";")]
prof = rt.velocity_profile()
g = rt.grid(C0, C1, (nn, nn, nn), neighbour_levels=1, coord_system="spherical")
n = g.n
X, Y, Z = g.coordinates()
U = torch.from_numpy(rt.interpolate_velocity(np.minimum(np.sqrt(X * X + Y * Y + Z * Z), R),
                                             rt.LinearInterpolation(prof.r, prof.Vp))).cuda()
del X, Y, Z
lat = int(np.ceil(np.sqrt(nsrc)))
srcs = np.array([1 + (nn * (2 * (q % lat) + 1)) // (2 * lat) + nn * ((nn * (2 * (q // lat) + 1)) // (2 * lat) + nn * (nn - 1))
                 for q in range(nsrc)], np.int64)
d = torch.empty((nsrc, n), dtype=torch.float64, device="cuda")
p = torch.empty((nsrc, n), dtype=torch.int32, device="cuda")
g._handle.set_option("schedule", 1)
ref = None
for b in modes:
    g._handle.set_option("batch", b)
    for rep in range(2):
        st = rt.RtStats()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rt.api.check(rt.lib().rt_bfm_solve_dev(g._handle.h, U.data_ptr(), srcs, nsrc, 64, d.data_ptr(), p.data_ptr(), C.byref(st)))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
    if ref is None:
        ref = (d.clone(), p.clone())
    sd = st.as_dict()
    print(json.dumps(dict(nn=nn, nsrc=nsrc, batch=b, ms_total=round(dt, 2), ms_per_source=round(dt / nsrc, 2),
                          kernel_ms=round(sd["kernel_ms"], 2), rounds=sd["sweeps"], launches=sd["total_launches"],
                          same_dist=bool(torch.equal(ref[0], d)), same_prev=bool(torch.equal(ref[1], p)))), flush=True)
