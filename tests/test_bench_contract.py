"""bench.py contract (CPU side): the reference arm prints ONE JSON line with the required keys; the workload table
names the BASELINE configs; ranks other than 0 of the reference arm exit without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          env=e, timeout=600)


def test_reference_arm_json_line():
    r = run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "annulus_180_50_20km"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "GTEPS" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["config"]["workload"] == "annulus_180_50_20km" and d["value"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    r = run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "annulus_180_50_20km"],
            env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_default_workload_is_config_1():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.DEFAULT_WORKLOAD == "annulus_1440_400_0.25km"
    w = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
    assert (w["ntheta"], w["nr"], w["spacing"]) == (1440, 400, 0.25)
