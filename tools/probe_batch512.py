"""GPU probe: BASELINE configs[2] on one GPU -- 512 sources on annulus 720x200 (default spacing), lock-step batch widths.
RT_PROBE_ONCE=1: a single 512-wide solve (for an ncu launch list)."""
import ctypes as C
import os
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
nsrc = int(os.environ.get("RT_PROBE_NSRC", "512"))
gr, G, halo = rt.init_annulus(720, 200, spacing=20.0, export=False)
h = gr._handle
n = gr.nnods
x_d, z_d, th_d, r_d = h.coords_dev()
U = torch.empty(n, dtype=torch.float64, device="cuda")
rt.api.check(rt.lib().rt_interp_velocity_dev(itp.knots, itp.values, len(itp.knots), r_d, n, -1.0, U.data_ptr()))
srcs = np.ascontiguousarray(rt.closest_point(gr, 2 * np.pi * np.arange(nsrc) / nsrc, np.full(nsrc, R), "polar"))
d = torch.empty((nsrc, n), dtype=torch.float64, device="cuda")
p = torch.empty((nsrc, n), dtype=torch.int32, device="cuda")
h.set_option("schedule", 1)
once = os.environ.get("RT_PROBE_ONCE") == "1"
ref = None
factors = [float(v) for v in os.environ.get("RT_DELTA_FACTORS", "0").split(",")]
for batch, fac in ([(512, factors[0])] if once else [(b, f) for f in factors for b in ((32, 128, 512) if len(factors) == 1 else (512,))]):
    h.set_option("batch", batch)
    h.set_option("delta_factor", fac)
    st = rt.RtStats()
    for rep in range(1 if once else 2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rt.api.check(rt.lib().rt_bfm_solve_dev(h.h, U.data_ptr(), srcs, nsrc, 64, d.data_ptr(), p.data_ptr(), C.byref(st)))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    if ref is None:
        ref = d.clone()
    sd = st.as_dict()
    print("delta_factor %g batch %3d: %.2f ms/source  (%.1f sources/s, rounds %d, launches %d, same dist %s)" %
          (fac, batch, dt / nsrc * 1e3, nsrc / dt, sd["sweeps"], sd["total_launches"], bool(torch.equal(ref, d))), flush=True)
