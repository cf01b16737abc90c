#!/usr/bin/env python
"""bench.py -- SSSP throughput of the B200-native bfm path (metric of BASELINE.json: relaxed edges/s in GTEPS
plus ms/source, next to the reference's CPU bfm).

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

A "step" = one pass of the hot path over one batch of synthetic input = ONE single-source bfm solve per GPU
(inputs resident in HBM).  With N > 1 (torchrun, one rank per GPU) every rank owns a full replica of the mesh
and solves its own shard of the source batch (weak scaling, no collective inside the relaxation); the
travel-time / predecessor tables of a step are then all-gathered over NCCL inside the timed region.

value   = TEPS_graph = E_graph * sources_solved / time   (Graph500-style: E_graph = sum of the reference's scan
          list lengths, schedule independent, so the ratio against the CPU arm is a pure time ratio)
roofline= HBM roofline of the relaxation kernel with the algorithmic bytes of SURVEY.md 8(d):
          12 B per relaxed candidate + B_v per active-vertex update (52 B 2-D / 60 B 3-D), over the relax
          kernel's own CUDA-event time on its launching stream.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R = 6371.0
WORKLOADS = {
    # BASELINE.json configs[3]: 3-D spherical-shell grid ~10M nodes, star1, single source, AK135
    "grid3d_216": dict(kind="3d", nn=(216, 216, 216), cpu_nn=(128, 128, 128), dim=3),
    # BASELINE.json configs[4]: 3-D grid ~50M nodes (368^3), multi-source batch
    "grid3d_368": dict(kind="3d", nn=(368, 368, 368), cpu_nn=(128, 128, 128), dim=3),
    "grid3d_128": dict(kind="3d", nn=(128, 128, 128), cpu_nn=(96, 96, 96), dim=3),
    "grid3d_64": dict(kind="3d", nn=(64, 64, 64), cpu_nn=(64, 64, 64), dim=3),
    # BASELINE.json configs[0]: README example, annulus 180x50, spacing 1 km
    "annulus_180_50_1km": dict(kind="2d", ntheta=180, nr=50, spacing=1.0, cpu=(180, 50, 20.0), dim=2),
    "annulus_180_50_5km": dict(kind="2d", ntheta=180, nr=50, spacing=5.0, cpu=(180, 50, 20.0), dim=2),
    "annulus_180_50_20km": dict(kind="2d", ntheta=180, nr=50, spacing=20.0, cpu=(180, 50, 20.0), dim=2),
    # BASELINE.json configs[1]: annulus 1440x400, spacing 0.25 km (~107M nodes)
    "annulus_1440_400_0.25km": dict(kind="2d", ntheta=1440, nr=400, spacing=0.25, cpu=(180, 50, 20.0), dim=2),
    # BASELINE.json configs[2] mesh: annulus 720x200 default spacing
    "annulus_720_200_20km": dict(kind="2d", ntheta=720, nr=200, spacing=20.0, cpu=(720, 200, 20.0), dim=2),
}
DEFAULT_WORKLOAD = os.environ.get("RT_BENCH_WORKLOAD", "annulus_1440_400_0.25km")
SHELL_C0 = (np.deg2rad(70.0), np.deg2rad(70.0), R - 2000.0)  # benchmarks/cpu.jl:9-13 rescaled to km
SHELL_C1 = (np.deg2rad(110.0), np.deg2rad(110.0), R)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.th = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for s in self.samples for k in range(4) if len(s) >= 7 and "Active" == s[3 + k]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def ncu_traffic(kernel, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` captures (profiles/traffic.json); None if that kernel / workload was not captured."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p))
    return t.get("%s@%s" % (kernel, workload))


def ak135():
    d = np.load(os.path.join(ROOT, "raytracer.jl_b200", "data", "ak135_profile.npz"))
    return (d["depth_km"].max() - d["depth_km"])[::-1].copy(), d["vp"][::-1].copy()


# ------------------------------------------------------------------------------------------------ GPU arm
class GpuWorkload:
    def __init__(self, name, rt, torch, schedule="near-far", sps=1):
        self.name, self.rt, self.torch = name, rt, torch
        self.w = WORKLOADS[name]
        prof = rt.velocity_profile()
        itp = rt.LinearInterpolation(prof.r, prof.Vp)
        if self.w["kind"] == "3d":
            self.g = rt.grid(SHELL_C0, SHELL_C1, self.w["nn"], neighbour_levels=1, coord_system="spherical")
            self.handle, self.n = self.g._handle, self.g.n
            X, Y, Z = self.g.coordinates()
            rr = np.minimum(np.sqrt(X * X + Y * Y + Z * Z), R)
            self.U_host = rt.interpolate_velocity(rr, itp)
            nx, ny, nz = self.w["nn"]
            # surface-centre node first, then a regular lattice of surface nodes for the source shards
            self.sources = [1 + (nx // 2) + nx * ((ny // 2) + ny * (nz - 1))]
            for a in range(8):
                for b in range(8):
                    self.sources.append(1 + (nx * (2 * a + 1)) // 16 + nx * ((ny * (2 * b + 1)) // 16 + ny * (nz - 1)))
            self.bv = 60
        else:
            gr, G, halo = rt.init_annulus(self.w["ntheta"], self.w["nr"], spacing=self.w["spacing"], export=False)
            self.gr = gr
            self.handle, self.n = gr._handle, gr.nnods
            x_d, z_d, th_d, r_d = self.handle.coords_dev()
            U = torch.empty(self.n, dtype=torch.float64, device="cuda")
            rt.api.check(rt.lib().rt_interp_velocity_dev(itp.knots, itp.values, len(itp.knots), r_d, self.n, -1.0,
                                                         U.data_ptr()))
            self.U_host = U.cpu().numpy()
            k = np.arange(65)
            self.sources = [int(s) for s in rt.closest_point(gr, 2 * np.pi * k / 65.0, np.full(65, R), "polar")]
            self.bv = 52
        self.sps = sps
        self.U_dev = torch.from_numpy(self.U_host).cuda()
        self.dist_dev = torch.empty(self.n * sps, dtype=torch.float64, device="cuda")
        self.prev_dev = torch.empty(self.n * sps, dtype=torch.int32, device="cuda")
        self.handle.set_option("profile_timers", 0)
        self.schedule = schedule
        self.handle.set_option("schedule", {"jacobi": 0, "near-far": 1}[schedule])
        # pinned host buffers of the end-to-end arm
        self.U_pin = torch.from_numpy(self.U_host).pin_memory()
        self.dist_pin = torch.empty(self.n * sps, dtype=torch.float64).pin_memory()
        self.prev_pin = torch.empty(self.n * sps, dtype=torch.int64).pin_memory()

    def solve_dev(self, source):
        import ctypes as C
        st = self.rt.RtStats()
        src = np.ascontiguousarray(np.atleast_1d(np.asarray(source, np.int64)))
        self.rt.api.check(self.rt.lib().rt_bfm_solve_dev(self.handle.h, self.U_dev.data_ptr(), src, len(src), 64,
                                                         self.dist_dev.data_ptr(), self.prev_dev.data_ptr(),
                                                         C.byref(st)))
        return st.as_dict()

    def solve_e2e(self, source):
        """The reference-facing call: host buffers in, host tables out (H2D of U, D2H of dist and prev inside)."""
        import ctypes as C
        st = self.rt.RtStats()
        src = np.ascontiguousarray(np.atleast_1d(np.asarray(source, np.int64)))
        self.rt.api.check(self.rt.lib().rt_bfm_solve(self.handle.h, self.U_pin.numpy(), src, len(src), 64,
                                                     self.dist_pin.data_ptr(), self.prev_pin.data_ptr(), C.byref(st)))
        return st.as_dict()


def run_gpu(args):
    import torch
    import rt_loader
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rt_loader.load_build().build()
    rt = rt_loader.load()
    rt.api.check(rt.lib().rt_set_device(local))
    wl = GpuWorkload(args.workload, rt, torch, args.schedule, args.sources_per_step)
    sps = args.sources_per_step
    n = wl.n
    gather_d = gather_p = None
    if dist_on:
        gather_d = torch.empty(world * n * sps, dtype=torch.float64, device="cuda")
        gather_p = torch.empty(world * n * sps, dtype=torch.int32, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()

    def pick(k):
        if sps == 1:
            return wl.sources[(rank + k * world) % len(wl.sources)] if dist_on else wl.sources[0]
        return [wl.sources[(rank + (k * sps + q) * world) % len(wl.sources)] for q in range(sps)]

    def step(k):
        st = wl.solve_dev(pick(k))
        if dist_on:  # gather the travel-time tables of this step on every rank
            dist.all_gather_into_tensor(gather_d, wl.dist_dev)
            dist.all_gather_into_tensor(gather_p, wl.prev_dev)
        return st

    for k in range(args.warmup):
        step(k)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    acc = dict(sweeps=0, relaxed_edges=0, vertex_updates=0, kernel_ms=0.0, relax_ms=0.0, relax_launches=0,
               total_launches=0)
    e_graph = 0
    for k in range(args.steps):
        st = step(args.warmup + k)
        for key in acc:
            acc[key] += st[key]
        e_graph = st["graph_edges"]
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop() if rank == 0 else None
    wall_ms = (t1 - t0) * 1e3
    times = torch.tensor([wall_ms, acc["kernel_ms"]], dtype=torch.float64, device="cuda")
    if dist_on:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    wall_ms, dev_ms = float(times[0]), float(times[1])

    # roofline pass: same step with per-launch CUDA-event timers around the relaxation kernel (the timers force a
    # host sync per round, so they stay out of the timed region above)
    wl.handle.set_option("profile_timers", 1)
    prof = dict(relaxed_edges=0, vertex_updates=0, relax_ms=0.0, relax_launches=0, kernel_ms=0.0, prev_ms=0.0)
    prof_steps = max(1, min(args.steps, 2))
    for k in range(prof_steps):
        st = wl.solve_dev(pick(0) if sps == 1 else wl.sources[0])
        for key in prof:
            prof[key] += st[key]
    wl.handle.set_option("profile_timers", 0)

    # end-to-end arm: same step through the host-buffer ABI call (H2D + D2H inside the timed region)
    e2e_steps = max(1, min(args.steps, 3))
    wl.solve_e2e(pick(0))
    barrier()
    t2 = time.perf_counter()
    for k in range(e2e_steps):
        wl.solve_e2e(pick(k))
    barrier()
    e2e_ms = torch.tensor([(time.perf_counter() - t2) * 1e3], dtype=torch.float64, device="cuda")
    if dist_on:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms[0])

    if rank == 0:
        peak, peak_src = peaks()
        nsolved = args.steps * world * sps
        value = e_graph * nsolved / (wall_ms * 1e-3) / 1e9
        kname = ("relax%s_kernel" if wl.schedule == "jacobi" else "push%s_kernel") % ("3d" if wl.w["kind"] == "3d" else "2d")
        bytes_alg = prof["relaxed_edges"] * 12 + prof["vertex_updates"] * wl.bv
        relax_ms = max(prof["relax_ms"], 1e-9)
        achieved = bytes_alg / (relax_ms * 1e-3) / 1e9
        out = {
            "metric": "sssp_relaxed_edges_per_s", "value": value, "unit": "GTEPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "nodes": n, "graph_edges_per_source": e_graph,
                       "sources_per_step_per_gpu": sps, "velocity": "AK135 Vp",
                       "schedule": "jacobi (reference sweeps)" if wl.schedule == "jacobi" else
                       "near-far push (dist bit-identical, prev exact except ties)",
                       "l2": "inputs larger than L2 (%.0f MB of node state per sweep set)" % (n * 48 / 1e6),
                       "parallelism": "source-sharded x%d, NCCL all_gather of tables" % world},
            "ms_per_source": wall_ms / args.steps / sps,
            "device_ms_per_step": dev_ms / args.steps,
            "teps_graph_gteps": value,
            "relax_rate_gteps": (bytes_alg / 12.0) / (relax_ms * 1e-3) / 1e9,
            "relaxed_edges_per_source": acc["relaxed_edges"] / args.steps,
            "sweeps_per_source": acc["sweeps"] / args.steps,
            "relax_kernel_share_of_step": relax_ms / max(prof["kernel_ms"], 1e-9),
            "prev_pass_share_of_step": prof["prev_ms"] / max(prof["kernel_ms"], 1e-9),
            "gpu_launches": acc["total_launches"],
            "clocks": clocks,
            "e2e": {"value": e_graph * e2e_steps * world * sps / (e2e_ms * 1e-3) / 1e9, "unit": "GTEPS",
                    "ms_per_source": e2e_ms / e2e_steps / sps, "h2d_bytes_per_step": n * 8 + 8 * sps,
                    "d2h_bytes_per_step": n * 16 * sps},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(kname, args.workload),
                         "peak_source": peak_src, "kernel": kname,
                         "bytes_model": "12 B per relaxed candidate + %d B per active-vertex update" % wl.bv,
                         "avg_launch_ms": relax_ms / max(prof["relax_launches"], 1)},
        }
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(args.workload, steps=1)
        print(json.dumps(out))
    if dist_on:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_instance(workload):
    from oracle import oracle as O
    w = WORKLOADS[workload]
    kr, kv = ak135()
    if w["kind"] == "3d":
        nn = w["cpu_nn"]
        X, Y, Z = O.grid3d_coords(SHELL_C0, SHELL_C1, nn, 1)
        rr = np.minimum(np.sqrt(X * X + Y * Y + Z * Z), R)
        U = O.interp_velocity(kr, kv, rr)
        src = 1 + (nn[0] // 2) + nn[0] * ((nn[1] // 2) + nn[1] * (nn[2] - 1))
        run = lambda th: O.bfm3d(nn, 1, X, Y, Z, U, src, nthreads=th)
        desc = "3-D shell %dx%dx%d star1 (same grid family, reduced), full single-source solve" % nn
    else:
        nt, nr, sp = w["cpu"]
        m = O.Annulus(nt, nr, sp)
        U = O.interp_velocity(kr, kv, m.r)
        src = O.closest_point(m.theta, m.r, 0.0, R)
        run = lambda th: O.bfm(m, U, src, nthreads=th)
        desc = "annulus %dx%d spacing %g km (reduced spacing), full single-source solve" % (nt, nr, sp)
    return run, desc


def host_threads():
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: ignore it, the CPU arm is
    meant to use the whole box)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(workload, steps=1):
    from oracle import oracle as O
    run, desc = cpu_instance(workload)
    cores = host_threads()
    t0 = time.perf_counter()
    for _ in range(steps):
        d, p, st = run(cores)
    dt = (time.perf_counter() - t0) / steps
    return {"value": st["graph_edges"] / dt / 1e9, "unit": "GTEPS", "cores": cores, "kind": "port",
            "sample": desc + "; TEPS_graph = E_graph / t; restated reference bfm (OpenMP, %d threads)" % cores,
            "ms_per_source": dt * 1e3, "relax_rate_gteps": st["relaxed_edges"] / dt / 1e9,
            "sweeps": st["sweeps"]}


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (oracle port; Julia is not installed and the
    reference has no compilable sources) on the host cores, same metric/unit, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    run, desc = cpu_instance(args.workload)
    cores = host_threads()
    for _ in range(args.warmup):
        run(cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        d, p, st = run(cores)
    dt = (time.perf_counter() - t0) / args.steps
    v = st["graph_edges"] / dt / 1e9
    out = {"impl": "reference", "metric": "sssp_relaxed_edges_per_s", "value": v, "unit": "GTEPS",
           "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": args.workload, "graph_edges_per_source": st["graph_edges"]},
           "cpu_baseline": {"value": v, "unit": "GTEPS", "cores": cores, "kind": "port", "sample": desc},
           "e2e": {"value": v, "unit": "GTEPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--schedule", default="near-far", choices=["jacobi", "near-far"])
    ap.add_argument("--sources-per-step", type=int, default=1,
                    help="sources solved per step and GPU (BASELINE config[2]: batches of earthquakes on one mesh)")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_gpu(a)
