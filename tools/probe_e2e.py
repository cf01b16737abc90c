"""GPU probe: where does the end-to-end (host buffers) time of rt_bfm_solve go?  3-D grid, one source."""
import ctypes as C
import sys
import time

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import numpy as np
import torch
import rt_loader

rt = rt_loader.load()
R = rt.R
nn = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "216,216,216").split(","))
g = rt.grid((np.deg2rad(70.0), np.deg2rad(70.0), R - 2000.0), (np.deg2rad(110.0), np.deg2rad(110.0), R), nn, 1, "spherical")
n = g.n
X, Y, Z = g.coordinates()
prof = rt.velocity_profile()
U = rt.interpolate_velocity(np.minimum(np.sqrt(X * X + Y * Y + Z * Z), R), rt.LinearInterpolation(prof.r, prof.Vp))
src = np.array([1 + nn[0] // 2 + nn[0] * (nn[1] // 2 + nn[1] * (nn[2] - 1))], np.int64)
h = g._handle
for sched in (0, 1):
    h.set_option("schedule", sched)
    for dfac in ((0,) if sched == 0 else (4, 8, 16, 32, 64)):
        h.set_option("delta_factor", dfac)
        U_pin = torch.from_numpy(U).pin_memory()
        d_pin = torch.empty(n, dtype=torch.float64).pin_memory()
        p_pin = torch.empty(n, dtype=torch.int64).pin_memory()
        U_dev = torch.from_numpy(U).cuda()
        d_dev = torch.empty(n, dtype=torch.float64, device="cuda")
        p_dev = torch.empty(n, dtype=torch.int32, device="cuda")
        st = rt.RtStats()

        def t(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps * 1e3

        dev = t(lambda: rt.api.check(rt.lib().rt_bfm_solve_dev(h.h, U_dev.data_ptr(), src, 1, 64, d_dev.data_ptr(), p_dev.data_ptr(), C.byref(st))))
        sd = st.as_dict()
        both = t(lambda: rt.api.check(rt.lib().rt_bfm_solve(h.h, U_pin.numpy(), src, 1, 64, d_pin.data_ptr(), p_pin.data_ptr(), C.byref(st))))
        donly = t(lambda: rt.api.check(rt.lib().rt_bfm_solve(h.h, U_pin.numpy(), src, 1, 64, d_pin.data_ptr(), None, C.byref(st))))
        none = t(lambda: rt.api.check(rt.lib().rt_bfm_solve(h.h, U_pin.numpy(), src, 1, 64, None, None, C.byref(st))))
        pag = t(lambda: rt.api.check(rt.lib().rt_bfm_solve(h.h, U, src, 1, 64, d_pin.data_ptr(), p_pin.data_ptr(), C.byref(st))))
        print("sched %d dfac %g: dev %.1f ms (rounds %d, ratio %.2f) | e2e both %.1f, dist only %.1f, no outputs %.1f, pageable U %.1f" %
              (sched, dfac, dev, sd["sweeps"], sd["relaxed_edges"] / sd["graph_edges"], both, donly, none, pag), flush=True)
