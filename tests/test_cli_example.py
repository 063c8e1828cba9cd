"""examples/jots_b200_run.cpp: the C++ host program over the C ABI (counterpart of jots_main.cpp)."""
import os
import subprocess

import pytest

from cases import make_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "jots_b200_run")
    lib = os.path.join(ROOT, "jots_b200")
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "jots_b200_run.cpp"),
                           "-L" + lib, "-ljots_b200", "-Wl,-rpath," + lib, "-pthread", "-o", exe])
    return exe


def test_cli_fails_loudly_without_a_gpu(tmp_path):
    import jots_b200 as J
    if J.load_library().jots_device_count() > 0:
        pytest.skip("a CUDA device is present")
    exe = _build(tmp_path)
    ini = make_case("NL_Reinert_B2", tmp_path)
    r = subprocess.run([exe, "-i", ini], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_cli_runs_an_unmodified_config(tmp_path):
    exe = _build(tmp_path)
    ini = make_case("NL_Reinert_B2", tmp_path, steps=20)
    r = subprocess.run([exe, "-i", ini], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Steps: 20" in r.stdout and "Newton iterations: 20" in r.stdout


@pytest.mark.gpu
def test_cli_multi_gpu_threads(tmp_path):
    """-g N: one host thread + one NCCL rank per GPU (the counterpart of mpirun -np N jots)."""
    import jots_b200 as J
    n = min(J.load_library().jots_device_count(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    exe = _build(tmp_path)
    ini = make_case("NL_Reinert_B2", tmp_path, steps=10)
    one = subprocess.run([exe, "-i", ini], capture_output=True, text=True)
    many = subprocess.run([exe, "-i", ini, "-g", str(n)], capture_output=True, text=True)
    assert one.returncode == 0 and many.returncode == 0, many.stderr
    t1 = [l for l in one.stdout.splitlines() if l.startswith("Temperature")][0]
    tn = [l for l in many.stdout.splitlines() if l.startswith("Temperature")][0]
    a = [float(x) for x in t1.replace("Temperature: min", "").replace("max", "").split()]
    b = [float(x) for x in tn.replace("Temperature: min", "").replace("max", "").split()]
    assert abs(a[0] - b[0]) < 1e-6 and abs(a[1] - b[1]) < 1e-6
    assert "DOFs: 1791" in many.stdout
