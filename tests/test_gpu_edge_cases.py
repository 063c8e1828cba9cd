"""Edge cases of the conduction solve through the C ABI: degenerate meshes, BC sets, polynomial limits,
solver failure semantics (Newton non-convergence is fatal, Krylov non-convergence is not:
/root/reference/src/nl_conduction_operator.cpp:312)."""
import numpy as np
import pytest

from cases import ini_text, write_brick_mesh
from test_gpu_parity import HF0, PROPS, FIELD_TOL, build, rel, rnd, smooth_state

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def J():
    import jots_b200
    if jots_b200.load_library().jots_device_count() < 1:
        pytest.fail("no CUDA device")
    return jots_b200


def _pair(J, tmp_path, txt):
    from oracle import jots as OJ
    ini = str(tmp_path / "c.ini")
    with open(ini, "w") as f:
        f.write(txt)
    return J.Driver(ini), OJ.JOTSDriver(OJ.Config(ini))


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_single_element_mesh(J, tmp_path, p):
    dev, orc = build(J, tmp_path, "Nonlinear_Unsteady", 3, p, "k", n=(1, 1, 1), steps=2)
    assert dev.run() == 2
    assert rel(dev.get_u(), orc.run()) < FIELD_TOL


def test_q4_hex_forms(J, tmp_path):
    dev, orc = build(J, tmp_path, "Nonlinear_Unsteady", 3, 4, "full", n=(2, 2, 1), grade=0.3)
    n = dev.n
    un, k, x = smooth_state(orc), 20.0 * rnd(n, 5), rnd(n, 6)
    orc.it.u_n = un
    orc.it.new_state(k)
    assert rel(dev.op.form_apply(J.FORM_RESIDUAL, k, np.empty(n), state=un), orc.it.mult(k)) < 1e-11
    Jm = orc.it.gradient(k)
    assert rel(dev.op.form_apply(J.FORM_SYSTEM, x, np.empty(n), state=un + orc.it.dt * k), Jm @ x) < 1e-11


def test_pure_neumann_and_all_dirichlet(J, tmp_path):
    mesh = str(tmp_path / "m.mesh")
    write_brick_mesh(mesh, 4, 3, 2, 0.01, 0.01, 0.01)
    for bcs in ([HF0, "HeatFlux, 1e5", HF0, "HeatFlux, -1e5", HF0, HF0],
                ["Isothermal, 300", "Isothermal, 310", "Isothermal, 320", "Isothermal, 330", "Isothermal, 340", "Isothermal, 350"]):
        txt = ini_text("Linearized_Unsteady", mesh, 300, PROPS["k"], bcs, steps=3, order=2)
        dev, orc = _pair(J, tmp_path, txt)
        assert dev.run() == 3
        assert rel(dev.get_u(), orc.run()) < FIELD_TOL
        # later Dirichlet BCs overwrite earlier ones on shared edges / corners (jots_driver.cpp:684-692)
        assert (dev.ctx.essential_dofs() == orc.it.ess).all()


def test_polynomial_limits(J, tmp_path):
    mesh = str(tmp_path / "m.mesh")
    write_brick_mesh(mesh, 2, 2, 2, 0.01, 0.01, 0.01)
    bcs = ["HeatFlux, 1e4", HF0, HF0, "Isothermal, 300", HF0, HF0]
    k8 = "Polynomial, 1e-20, -1e-17, 1e-14, -1e-11, 1e-8, 1e-5, 0.01, 5"        # 8 coefficients: the maximum
    props = [("Density_rho", "Uniform, 8000"), ("Specific_Heat_C", "Uniform, 500"), ("Thermal_Conductivity_k", k8)]
    dev, orc = _pair(J, tmp_path, ini_text("Nonlinear_Unsteady", mesh, 300, props, bcs, newton=50, steps=2, order=2))
    assert dev.run() == 2
    assert rel(dev.get_u(), orc.run()) < FIELD_TOL
    props[2] = ("Thermal_Conductivity_k", k8 + ", 1")
    ini = str(tmp_path / "bad.ini")
    with open(ini, "w") as f:
        f.write(ini_text("Nonlinear_Unsteady", mesh, 300, props, bcs, newton=50, steps=2, order=2))
    with pytest.raises(J.JotsError, match="more than 8"):
        J.Driver(ini)
    props[2] = ("Thermal_Conductivity_k", "Polynomial, 5")                        # material_property.cpp:70
    with open(ini, "w") as f:
        f.write(ini_text("Nonlinear_Unsteady", mesh, 300, props, bcs, newton=50, steps=2, order=2))
    with pytest.raises(J.JotsError, match=">= 2"):
        J.Driver(ini)


def test_config_errors(J, tmp_path):
    mesh = str(tmp_path / "m.mesh")
    write_brick_mesh(mesh, 2, 2, 2, 0.01, 0.01, 0.01)
    ini = str(tmp_path / "c.ini")
    with open(ini, "w") as f:   # 5 BCs for 6 boundary attributes
        f.write(ini_text("Linearized_Unsteady", mesh, 300, PROPS["const"], [HF0] * 5, steps=1))
    with pytest.raises(J.JotsError, match="BC count"):
        J.Driver(ini)
    with open(ini, "w") as f:
        f.write(ini_text("Linearized_Unsteady", mesh, 300, PROPS["const"], [HF0] * 5 + ["Robin, 1"], steps=1))
    with pytest.raises(J.JotsError, match="Invalid/Unknown boundary condition"):
        J.Driver(ini)
    with open(ini, "w") as f:
        f.write(ini_text("Linearized_Unsteady", mesh, 300, PROPS["const"], [HF0] * 6, steps=1, solver="BiCGSTAB"))
    with pytest.raises(J.JotsError, match="solver"):
        J.Driver(ini)


def test_newton_non_convergence_is_fatal_krylov_is_not(J, tmp_path):
    dev, orc = build(J, tmp_path, "Nonlinear_Unsteady", 3, 2, "full", newton=1, steps=1, n=(4, 3, 2))
    with pytest.raises(J.JotsError, match="Newton solver did not converge"):
        dev.run()
    # 2 Krylov iterations per solve: not converged, not an error; same field as the oracle doing the same
    mesh = str(tmp_path / "m2.mesh")
    write_brick_mesh(mesh, 4, 3, 2, 0.01, 0.01, 0.01)
    txt = ini_text("Linearized_Unsteady", mesh, 300, PROPS["k"], ["HeatFlux, 7.5e5", HF0, HF0, "Isothermal, 300", HF0, HF0], steps=2)
    txt = txt.replace("Max_Iterations=100", "Max_Iterations=2", 1)
    dev, orc = _pair(J, tmp_path, txt)
    assert dev.run() == 2
    assert dev.ctx.stats()["last_linear_converged"] == 0
    assert rel(dev.get_u(), orc.run()) < 1e-9


def test_zero_time_steps(J, tmp_path):
    dev, orc = build(J, tmp_path, "Linearized_Unsteady", 3, 1, "const", steps=0)
    assert dev.run() == 0
    assert np.all(dev.get_u() == 300.0)
