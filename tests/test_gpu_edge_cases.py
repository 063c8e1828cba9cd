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


@pytest.mark.parametrize("dim,order", [(2, 1), (3, 2)])
def test_restart_files_round_trip_and_continuation(J, tmp_path, dim, order):
    """The reference's Restart_Files regression (tests/linearized_unsteady_heat/Restart_Files.cpp:9-80: 2-D bar, Q1,
    Sinusoidal_Isothermal, restart written at the last cycle, re-loaded field within 1e-12) plus a 3-D Q2 case, and the
    stronger statement that N steps + restart + N steps equal 2N uninterrupted steps."""
    from cases import ini_text, write_quad_mesh, write_brick_mesh
    from test_gpu_parity import rel
    mesh = str(tmp_path / "m.mesh")
    if dim == 2:
        write_quad_mesh(mesh, 20, 3, 1.0, 0.1, attrs=(1, 3, 1, 2))
        props = [("Density_rho", "Uniform, 0.1"), ("Specific_Heat_C", "Uniform, 1000"), ("Thermal_Conductivity_k", "Uniform, 100")]
        bcs = ["HeatFlux, 0", "Isothermal, 300", "Sinusoidal_Isothermal, 5, 6.283185307, 0, 300"]
        dt = "0.001"
    else:
        write_brick_mesh(mesh, 5, 3, 2, 0.01, 0.02, 0.015, grade=0.3)
        props = [("Density_rho", "Uniform, 8000"), ("Specific_Heat_C", "Polynomial, 4.5, -850"), ("Thermal_Conductivity_k", "Polynomial, 0.09, -17")]
        bcs = ["HeatFlux, 7.5e5", "Isothermal, 350", "HeatFlux, 0", "Sinusoidal_Isothermal, 50, 40, 0.3, 400", "HeatFlux, 0", "HeatFlux, 0"]
        dt = "1e-2"
    n = 6
    prefix = str(tmp_path / "restart")

    def write(name, steps, restart_from=None, freq=0):
        txt = ini_text("Linearized_Unsteady", mesh, 300, props, bcs, steps=steps, dt=dt, order=order)
        txt = txt.replace("Restart_Freq=0", "Restart_Freq=%d" % freq)
        txt = txt.replace("Use_Restart=No", "Use_Restart=%s\nRestart_Prefix=%s\nRestart_Cycle=%d" % ("Yes" if restart_from is not None else "No", prefix, restart_from or 0))
        path = str(tmp_path / name)
        open(path, "w").write(txt)
        return path

    d0 = J.Driver(write("first.ini", n, freq=n))
    assert d0.run() == n
    u_n = d0.get_u()
    dr = J.Driver(write("restarted.ini", 2 * n, restart_from=n))
    assert abs(dr.time - d0.time) < 1e-15 and dr.n == d0.n
    assert np.abs(dr.get_u() - u_n).max() <= 1e-12 * np.abs(u_n).max()       # Restart_Files.cpp: EPSILON = 1e-12
    assert dr.run() == n                                                       # cycles n+1 .. 2n
    d2 = J.Driver(write("straight.ini", 2 * n))
    assert d2.run() == 2 * n
    assert rel(dr.get_u(), d2.get_u()) < 1e-10
    with pytest.raises(J.JotsError, match="restart"):
        J.Driver(write("missing.ini", n, restart_from=999))


@pytest.mark.parametrize("sim,order", [("Nonlinear_Unsteady", 2), ("Linearized_Unsteady", 3), ("Steady", 2)])
def test_paraview_output_at_the_visualization_frequency(J, tmp_path, monkeypatch, sim, order):
    """OutputManager::WriteVizOutput through the driver shim (jots_driver.cpp:571-573, 750-800): the initial condition and every
    Visualization_Freq-th cycle of a time-integrating run (one output at the end of a steady run), fields Rank, Temperature
    and one per material property of the input file -- the property evaluated at the field that is written."""
    from cases import ini_text, write_brick_mesh
    from oracle import vtkfmt
    out = str(tmp_path / "ParaView")
    monkeypatch.setenv("JOTS_PARAVIEW_DIR", out)
    mesh = str(tmp_path / "m.mesh")
    write_brick_mesh(mesh, 5, 3, 2, 0.01, 0.02, 0.015, grade=0.3)
    props = [("Density_rho", "Uniform, 8000"), ("Specific_Heat_C", "Polynomial, 4.5, -850"), ("Thermal_Conductivity_k", "Polynomial, 0.09, -17")]
    bcs = ["HeatFlux, 7.5e5", "Isothermal, 350", "HeatFlux, 0", "Isothermal, 400", "HeatFlux, 0", "HeatFlux, 0"]
    steady = sim == "Steady"
    txt = ini_text(sim, mesh, 300, props if not steady else [props[2]], bcs, time=not steady, steps=4, dt="1e-2", order=order,
                   newton=None if sim == "Linearized_Unsteady" else 50)
    txt = txt.replace("Visualization_Freq=0", "Visualization_Freq=2")
    ini = str(tmp_path / "c.ini")
    open(ini, "w").write(txt)
    dev = J.Driver(ini)
    dev.run()
    entries = vtkfmt.read_pvd(out + "/ParaView.pvd")
    if steady:
        assert len(entries) == 1
    else:
        assert [f for _, f in entries] == ["Cycle000000/data.pvtu", "Cycle000002/data.pvtu", "Cycle000004/data.pvtu"]
        assert np.allclose([t for t, _ in entries], [0.0, 0.02, 0.04], atol=1e-15)
    names = ["Rank", "Temperature"] + ([p[0] for p in props] if not steady else [props[2][0]])
    last = out + "/" + entries[-1][1].split("/")[0]
    assert vtkfmt.read_pvtu(last + "/data.pvtu")["point_arrays"] == names
    v = vtkfmt.read_vtu(last + "/proc000000.vtu")
    assert sorted(v["point_data"]) == sorted(names)
    s = dev.space()
    g, u = s.gather(), dev.get_u()
    nloc = (order + 1) ** 3
    T = v["point_data"]["Temperature"].reshape(-1, nloc)
    if order <= 2:
        assert np.array_equal(T, u[g])                     # Gauss-Lobatto nodes = uniform lattice: the nodal values themselves
    else:
        # corners are nodes of both lattices; all values stay within the nodal range up to interpolation overshoot
        corners = [0, order, order * (order + 1), (order + 1) ** 2 - 1]
        assert np.allclose(T[:, corners], u[g][:, corners], rtol=1e-13)
    k = v["point_data"]["Thermal_Conductivity_k"].reshape(-1, nloc)
    if order <= 2:
        assert np.allclose(k, 0.09 * u[g] - 17.0, rtol=1e-13)
    assert np.abs(v["point_data"]["Rank"]).max() == 0.0
    if not steady:
        v0 = vtkfmt.read_vtu(out + "/Cycle000000/proc000000.vtu")
        assert np.abs(v0["point_data"]["Temperature"] - 300.0).max() < 1e-12      # the initial condition, before any BC is applied
        assert np.allclose(v0["point_data"]["Density_rho"], 8000.0) and np.allclose(v0["point_data"]["Specific_Heat_C"], 4.5 * 300.0 - 850.0)
