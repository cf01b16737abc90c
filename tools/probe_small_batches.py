"""GPU probe: small batches (1 .. 64 sources) on the configs[2] mesh, persistent kernel against the launch sequence."""
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
gr, G, halo = rt.init_annulus(720, 200, spacing=20.0, export=False)
h = gr._handle
n = gr.nnods
x_d, z_d, th_d, r_d = h.coords_dev()
U = torch.empty(n, dtype=torch.float64, device="cuda")
rt.api.check(rt.lib().rt_interp_velocity_dev(itp.knots, itp.values, len(itp.knots), r_d, n, -1.0, U.data_ptr()))
h.set_option("schedule", 1)
for nsrc in (1, 4, 32, 64):
    srcs = np.ascontiguousarray(rt.closest_point(gr, 2 * np.pi * np.arange(nsrc) / nsrc, np.full(nsrc, R), "polar"))
    d = torch.empty((nsrc, n), dtype=torch.float64, device="cuda")
    p = torch.empty((nsrc, n), dtype=torch.int32, device="cuda")
    for pers in (1, 0):
        h.set_option("persistent", pers)
        st = rt.RtStats()
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rt.api.check(rt.lib().rt_bfm_solve_dev(h.h, U.data_ptr(), srcs, nsrc, 64, d.data_ptr(), p.data_ptr(), C.byref(st)))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        print("nsrc %3d persistent %d: %.2f ms/source (total %.1f ms, rounds %d)" % (nsrc, pers, dt / nsrc * 1e3, dt * 1e3, st.sweeps), flush=True)
