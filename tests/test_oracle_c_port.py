"""The C/OpenMP restatement used as the timed CPU baseline (oracle/c/jots_cpu_port.c) against the
NumPy oracle on the bench workload: same temperature field (<= 1e-10 relative l2) and identical
Newton / Krylov iteration counts."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_build", "jots_cpu_port")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "c")], stdout=subprocess.DEVNULL)


@pytest.mark.parametrize("n,p,kspec", [((5, 4, 3), 2, "Polynomial, 0.09, -17"), ((4, 4, 4), 1, "Polynomial, 0.09, -17"),
                                       ((3, 3, 2), 3, "Uniform, 10")])
def test_c_port_matches_numpy_oracle(tmp_path, n, p, kspec):
    from oracle import jots as OJ, mesh as OM
    from cases import ini_text
    _build()
    props = [("Density_rho", "Uniform, 8000"), ("Specific_Heat_C", "Uniform, 500"), ("Thermal_Conductivity_k", kspec)]
    bcs = ["HeatFlux, 7.5e5", "HeatFlux, 0", "HeatFlux, 0", "Isothermal, 300", "HeatFlux, 0", "HeatFlux, 0"]
    ini = str(tmp_path / "c.ini")
    with open(ini, "w") as f:
        f.write(ini_text("Nonlinear_Unsteady", "unused", 300, props, bcs, newton=100, solver="CG", prec="Jacobi", steps=3, order=p))
    lx, ly, lz = 0.01, 0.008, 0.006
    orc = OJ.JOTSDriver(OJ.Config(ini), mesh=OM.make_brick(*n, lx, ly, lz))
    u_ref = orc.run()
    kc = [t.strip() for t in kspec.split(",")[1:]]
    out = str(tmp_path / "u.bin")
    cmd = [BIN, *map(str, n), str(lx), str(ly), str(lz), str(p), "1e-2", "3", "8000", "500", "300", "7.5e5", "300",
           "1e-10", "1e-16", "100", "1e-10", "100", "2", str(len(kc)), *kc, "--out", out]
    res = json.loads(subprocess.check_output(cmd).decode())
    u = np.fromfile(out)
    assert u.size == u_ref.size == res["ndof"]
    assert np.linalg.norm(u - u_ref) / np.linalg.norm(u_ref) < 1e-10
    st = orc.it.stats
    assert res["newton_iters"] == st["newton_iters"] and res["residual_evals"] == st["residual_evals"]
    assert abs(res["krylov_iters"] - st["krylov_iters"]) <= 1
