"""bench.py's driver-facing contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`: the
C/OpenMP restatement of the reference's algorithm on the host cores) prints one JSON line with the agreed keys and says which
bounded sample it timed; ranks other than 0 print nothing; the GPU arm refuses to run without a device (no CPU fallback);
workload labels follow BASELINE.json's configs."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def run(args, env=None):
    e = dict(os.environ)
    e.pop("RANK", None)
    if env:
        e.update(env)
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=300, env=e, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-n", "8"])
    assert r.returncode == 0, r.stderr[-800:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GDOF/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["metric"].startswith("Krylov GDOF/s") and d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["vs_baseline"] is None and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "8^3" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    w = d["config"]["workload"]
    assert w.startswith("configs[3]: synthetic refined brick, 128x128x128 Q2 hexes, 16974593 DOFs")
    assert "TIMED HERE: a bounded sample of it, 8x8x8 Q2 hexes = 4913 DOFs" in w         # the label names what was really timed
    assert d["config"]["sample_n_elem"] == [8, 8, 8] and d["config"]["sample_n_dof"] == 17 ** 3


def test_reference_arm_is_silent_on_other_ranks():
    r = run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-n", "8", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_device():
    import jots_b200
    if jots_b200.load_library().jots_device_count() > 0:
        import pytest
        pytest.skip("a CUDA device is present")
    r = run(["--steps", "1", "--warmup", "0", "--no-cpu", "--no-strong"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_workload_labels_follow_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    a = argparse.Namespace(n=128, order=2, scaling="weak", variant="A", solver="CG", prec="Jacobi")
    c = bench.workload_config(a, 1)
    assert c["workload"].startswith("configs[3]:") and c["n_dof"] == 257 ** 3 and c["parallelism"] == "single GPU"
    c8 = bench.workload_config(a, 8)
    assert c8["n_elem"] == [256, 256, 256] and c8["n_dof"] == 513 ** 3 and "2x2x2" in c8["parallelism"]
    a3 = argparse.Namespace(n=84, order=3, scaling="weak", variant="A", solver="CG", prec="Jacobi")
    c3 = bench.workload_config(a3, 8)
    assert c3["workload"].startswith("configs[4]:") and c3["n_dof"] == 505 ** 3 == 128787625
    assert bench.box_dims(2) == [2, 1, 1] and bench.box_dims(4) == [2, 2, 1] and bench.box_dims(8) == [2, 2, 2]
