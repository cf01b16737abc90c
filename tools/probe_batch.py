"""GPU probe: multi-source throughput on the config[2] mesh (annulus 720x200, default spacing) -- sources per second
for several lock-step batch widths."""
import ctypes as C
import sys
import time

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import numpy as np
import torch
import rt_loader

rt = rt_loader.load()
R = rt.R
prof = rt.velocity_profile()
itp = rt.LinearInterpolation(prof.r, prof.Vp)
nt, nr, sp = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (720, 200, 20.0)
nsrc = int(sys.argv[4]) if len(sys.argv) > 4 else 64
gr, G, halo = rt.init_annulus(nt, nr, spacing=sp, export=False)
h = gr._handle
n = gr.nnods
x_d, z_d, th_d, r_d = h.coords_dev()
U = torch.empty(n, dtype=torch.float64, device="cuda")
rt.api.check(rt.lib().rt_interp_velocity_dev(itp.knots, itp.values, len(itp.knots), r_d, n, -1.0, U.data_ptr()))
srcs = np.ascontiguousarray(rt.closest_point(gr, 2 * np.pi * np.arange(nsrc) / nsrc, np.full(nsrc, R), "polar"))
d = torch.empty((nsrc, n), dtype=torch.float64, device="cuda")
p = torch.empty((nsrc, n), dtype=torch.int32, device="cuda")
h.set_option("schedule", 1)
import os
h.set_option("persistent", int(os.environ.get("RT_PERSISTENT", "-1")))
ref = None
for batch in (1, 4, 8, 16, 32):
    h.set_option("batch", batch)
    st = rt.RtStats()
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rt.api.check(rt.lib().rt_bfm_solve_dev(h.h, U.data_ptr(), srcs, nsrc, 64, d.data_ptr(), p.data_ptr(), C.byref(st)))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    if ref is None:
        ref = d.clone()
    sd = st.as_dict()
    print("batch %2d: %.1f ms/source  (%.1f sources/s, rounds %d, launches %d, same dist %s)" %
          (batch, dt / nsrc * 1e3, nsrc / dt, sd["sweeps"], sd["total_launches"], bool(torch.equal(ref, d))), flush=True)
