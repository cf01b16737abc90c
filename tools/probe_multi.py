"""GPU probe: rt_bfm_solve_multi (one process, one mesh replica + host thread per GPU) on the config[2] mesh.
usage: python tools/probe_multi.py [nsrc]"""
import json
import sys
import time

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import numpy as np
import rt_loader

rt = rt_loader.load()
nsrc = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ndev = rt.device_count()
prof = rt.velocity_profile()
itp = rt.LinearInterpolation(prof.r, prof.Vp)
reps = []
for d in range(ndev):
    rt.set_device(d)
    gr, G, halo = rt.init_annulus(720, 200, export=(d == 0))
    reps.append(gr)
rt.set_device(0)
Vp = rt.interpolate_velocity(reps[0].r, itp)
srcs = rt.closest_point(reps[0], 2 * np.pi * np.arange(nsrc) / nsrc, np.full(nsrc, rt.R), "polar")
ref = None
for k in sorted({1, ndev}):
    for rep in range(2):
        t = time.time()
        D = rt.bfm_multi(reps[:k], srcs, Vp, schedule="near-far")
        dt = time.time() - t
    if ref is None:
        ref = D.dist
    print(json.dumps(dict(replicas=k, nsrc=nsrc, n=int(reps[0].nnods), wall_s=round(dt, 3),
                          ms_per_source=round(1e3 * dt / nsrc, 2), kernel_ms_max=round(D.stats["kernel_ms"], 1),
                          same=bool(np.array_equal(ref, D.dist)))), flush=True)
