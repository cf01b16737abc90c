"""GPU probe: solve time / work of the 2-D annulus solver for several meshes, schedules and bucket widths.
usage: python tools/probe_annulus.py "nt,nr,spacing;..." "schedule:delta_factor;..."   (schedule 0|1)"""
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
prof = rt.velocity_profile()
itp = rt.LinearInterpolation(prof.r, prof.Vp)
meshes = [tuple(float(v) for v in m.split(",")) for m in (sys.argv[1] if len(sys.argv) > 1 else "180,50,20").split(";")]
modes = [tuple(float(v) for v in m.split(":")) for m in (sys.argv[2] if len(sys.argv) > 2 else "0:0;1:8").split(";")]
for (nt, nr, sp) in meshes:
    t = time.time()
    gr, G, halo = rt.init_annulus(int(nt), int(nr), spacing=sp, export=False)
    tb = time.time() - t
    h = gr._handle
    n = gr.nnods
    x_d, z_d, th_d, r_d = h.coords_dev()
    U = torch.empty(n, dtype=torch.float64, device="cuda")
    rt.api.check(rt.lib().rt_interp_velocity_dev(itp.knots, itp.values, len(itp.knots), r_d, n, -1.0, U.data_ptr()))
    src = rt.closest_point(gr, 0.0, R, "polar")
    import os
    h.set_option("profile_timers", int(os.environ.get("RT_TIMERS", "1")))
    h.set_option("persistent", int(os.environ.get("RT_PERSISTENT", "-1")))
    h.set_option("cta_units", int(os.environ.get("RT_CTA_UNITS", "0")))
    h.set_option("check_every", int(os.environ.get("RT_CHECK_EVERY", "0")))
    h.set_option("group_screen", int(os.environ.get("RT_GROUP_SCREEN", "1")))
    h.set_option("compact", int(os.environ.get("RT_COMPACT", "1")))
    h.set_option("use_graph", int(os.environ.get("RT_GRAPH", "1")))
    h.set_option("fuse_begin", int(os.environ.get("RT_FUSE", "1")))
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    p = torch.empty(n, dtype=torch.int32, device="cuda")
    ref = None
    for (sched, df) in modes:
        h.set_option("schedule", sched)
        if df < 0:
            h.set_option("delta", -df)   # negative: absolute bucket width in seconds
        else:
            h.set_option("delta", 0.0)
            h.set_option("delta_factor", df)
        st = rt.RtStats()
        s = np.array([src], np.int64)
        for rep in range(2):
            t = time.time()
            rt.api.check(rt.lib().rt_bfm_solve_dev(h.h, U.data_ptr(), s, 1, 64, d.data_ptr(), p.data_ptr(), C.byref(st)))
            ts = time.time() - t
        sd = st.as_dict()
        if ref is None:
            ref = d.clone()
        print(json.dumps(dict(cfg=(nt, nr, sp), n=n, build_s=round(tb, 2), sched=sched, dfac=df, solve_s=round(ts, 4),
                              rounds=sd["sweeps"], relaxed=sd["relaxed_edges"], ratio=round(sd["relaxed_edges"] / sd["graph_edges"], 2),
                              screened=sd["screened_edges"], exact=sd["exact_edges"], kernel_ms=round(sd["kernel_ms"], 1), relax_ms=round(sd["relax_ms"], 1), prev_ms=round(sd["prev_ms"], 1), launches=sd["total_launches"],
                              teps_graph_G=round(sd["graph_edges"] / ts / 1e9, 2), same_dist=bool(torch.equal(ref, d)))), flush=True)
    del gr, G, h
