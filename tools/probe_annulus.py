import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, rt_loader
rt = rt_loader.load()
R = rt.R
prof = rt.velocity_profile()
itp = rt.LinearInterpolation(prof.r, prof.Vp)
import ctypes as C
for (nt, nr, sp) in [(180, 50, 20.0), (180, 50, 5.0), (180, 50, 1.0), (720, 200, 20.0)]:
    t = time.time(); gr, G, halo = rt.init_annulus(nt, nr, spacing=sp, export=False); tb = time.time() - t
    h = gr._handle; n = gr.nnods
    import torch
    x_d, z_d, th_d, r_d = h.coords_dev()
    U = torch.empty(n, dtype=torch.float64, device='cuda')
    rt.api.check(rt.lib().rt_interp_velocity_dev(itp.knots, itp.values, len(itp.knots), r_d, n, -1.0, U.data_ptr()))
    src = rt.closest_point(gr, 0.0, R, 'polar')
    h.set_option('profile_timers', 1)
    st = rt.RtStats(); s = np.array([src], np.int64)
    d = torch.empty(n, dtype=torch.float64, device='cuda'); p = torch.empty(n, dtype=torch.int32, device='cuda')
    for rep in range(2):
        t = time.time()
        rt.api.check(rt.lib().rt_bfm_solve_dev(h.h, U.data_ptr(), s, 1, 64, d.data_ptr(), p.data_ptr(), C.byref(st)))
        ts = time.time() - t
    sd = st.as_dict()
    print(json.dumps(dict(cfg=(nt, nr, sp), n=n, sizes=[int(v) for v in h.sizes()], build_s=round(tb, 2), solve_s=round(ts, 3),
          sweeps=sd['sweeps'], relaxed=sd['relaxed_edges'], E_graph=sd['graph_edges'], ratio=round(sd['relaxed_edges'] / sd['graph_edges'], 1),
          relax_ms=round(sd['relax_ms'], 1), kernel_ms=round(sd['kernel_ms'], 1), relax_rate_G=round(sd['relaxed_edges'] / sd['relax_ms'] / 1e6, 1),
          teps_graph_G=round(sd['graph_edges'] / ts / 1e9, 2), T180=float(d[rt.closest_point(gr, float(np.deg2rad(np.float32(180.0))), R, 'polar') - 1]))), flush=True)
    del gr, G, h
