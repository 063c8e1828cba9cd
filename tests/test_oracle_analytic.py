"""Pins the CPU oracle on every regression the reference ships for the conduction solve
(/root/reference/tests/{linearized_unsteady_heat,nonlinear_unsteady_heat,steady_heat}/*.cpp:5)
with the reference's own error functional (tests/test_helper.hpp.in:134)."""
import os

import numpy as np
import pytest

from oracle import analytic, jots as OJ
from cases import CASES, REF, make_case


def _run(name, tmp_path, use_ref):
    ini = make_case(name, tmp_path, use_reference_mesh=use_ref)
    cfg = OJ.Config(ini)
    drv = OJ.JOTSDriver(cfg)
    u = drv.run()
    t_end = cfg.max_timesteps * cfg.dt if cfg.using_time_integration else 0.0
    err = analytic.reference_error(drv.space, u, getattr(analytic, CASES[name]["exact"]), t_end)
    return err, drv


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_analytic_pin(name, tmp_path):
    err, drv = _run(name, tmp_path, use_ref=False)
    assert np.isfinite(err) and err <= CASES[name]["eps"], (name, err)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name", ["Reinert_B2", "NL_Reinert_B2", "Danish_Model1_Neumann"])
def test_oracle_on_shipped_mesh_files(name, tmp_path):
    """Same pins on the reference's own mesh files (Pointwise vertex order, un-structured numbering)."""
    err, drv = _run(name, tmp_path, use_ref=True)
    assert np.isfinite(err) and err <= CASES[name]["eps"], (name, err)


@pytest.mark.parametrize("name", ["Reinert_B1", "Danish_Model1_Neumann"])
def test_committed_golden_fields_are_the_oracle_output(name, tmp_path):
    """tests/golden/regressions.npz (the fixture the GPU tests compare against) is reproducible."""
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "regressions.npz"))
    ini = make_case(name, tmp_path)
    u = OJ.JOTSDriver(OJ.Config(ini)).run()
    assert np.linalg.norm(u - gold[name]) <= 1e-12 * np.linalg.norm(gold[name])


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_oracle_convergence_order_on_a_smooth_nonlinear_steady_problem(p, tmp_path):
    """The reference's regressions pin the discretisation at Q2 only.  Here: steady conduction with
    k(T) = 1 + T, T(0) = 0, T(1) = 1 on [0,1] x [0,.2] x [0,.1] -- Kirchhoff transform theta = T + T^2/2 = 1.5 x, so
    T = -1 + sqrt(1 + 3x) is smooth -- must converge with order p+1 in the reference's error functional for every order
    the space supports (basis, Gauss-Lobatto nodes, p+2-point quadrature and the Newton Jacobian all enter)."""
    from oracle import mesh as OM
    exact = lambda x, t: -1.0 + np.sqrt(1.0 + 3.0 * x[:, 0])
    hf0 = "HeatFlux, 0"
    errs = []
    for nx in (2, 4, 8):
        from cases import ini_text
        ini = str(tmp_path / ("c%d.ini" % nx))
        with open(ini, "w") as f:
            f.write(ini_text("Steady", "unused", 0.5, [("Thermal_Conductivity_k", "Polynomial, 1, 1")],
                             ["Isothermal, 0", hf0, hf0, "Isothermal, 1", hf0, hf0], time=False, newton=200, solver="CG", order=p))
        drv = OJ.JOTSDriver(OJ.Config(ini), mesh=OM.make_brick(nx, 2, 1, 1.0, 0.2, 0.1))
        errs.append(analytic.reference_error(drv.space, drv.run(), exact, 0.0))
    orders = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
    assert p + 0.5 <= orders[0] <= p + 1.7 and p + 0.7 <= orders[1] <= p + 1.7, (p, errs, orders)


@pytest.mark.parametrize("p", [1, 2, 3])
def test_oracle_precice_boundary_data_path_is_exact_on_polynomials(tmp_path, p):
    """Pins for the oracle's restatement of the preCICE boundary data path (boundary_condition.cpp:77-206), which the
    reference only tests through preCICE itself: the wall heat flux -k(T) grad T . n of a field in the FE space is exact at
    the nodes (T linear, affine graded mesh: q = -k(T) dT/dn), and the Neumann vector of a nodal flux field integrates it
    exactly (sum_i b_i = int g dS for g in the trace space)."""
    from cases import ini_text, write_brick_mesh
    from oracle import jots as OJ
    props = [("Density_rho", "Uniform, 8000"), ("Specific_Heat_C", "Uniform, 500"), ("Thermal_Conductivity_k", "Polynomial, 0.09, -17")]
    hf0 = "HeatFlux, 0"
    mesh = str(tmp_path / "m.mesh")
    write_brick_mesh(mesh, 3, 2, 2, 0.01, 0.02, 0.015, grade=0.4)
    ini = str(tmp_path / "c.ini")
    with open(ini, "w") as f:
        f.write(ini_text("Nonlinear_Unsteady", mesh, 300, props, ["preCICE_Isothermal, M, 320", hf0, hf0, "preCICE_HeatFlux, M, 0", hf0, hf0],
                         newton=50, steps=1, order=p))
    orc = OJ.JOTSDriver(OJ.Config(ini))
    X = orc.space.node_coordinates()
    u = 300.0 + 1000.0 * X[:, 0] + 500.0 * X[:, 1] - 200.0 * X[:, 2]
    for attr, dTdn in ((1, -1000.0), (4, 1000.0)):
        bc = OJ.BoundaryCondition(attr, ["preCICE_Isothermal", "M", "1"])
        dofs, coords = orc.space.bdr_nodes(attr)
        assert np.allclose(coords[:, 0], 0.0 if attr == 1 else 0.01)
        q = bc.write_data(orc.space, u, orc.props["k"])
        assert np.abs(q + (0.09 * u[dofs] - 17.0) * dTdn).max() <= 1e-12 * np.abs(q).max()
    # nodal flux field on attribute 4 (x = 0.01, area 0.02 x 0.015): g = 1 + 100 y  ->  int g dS = A (1 + 100 * 0.01)
    bc = orc.bcs[3]
    dofs, coords = orc.space.bdr_nodes(4)
    bc.set_read_data(orc.space, 1.0 + 100.0 * coords[:, 1])
    b = orc.space.neumann_vector({4: bc.value})
    assert abs(b.sum() - 0.02 * 0.015 * 2.0) <= 1e-13
    assert np.abs(b[np.setdiff1d(np.arange(orc.space.ndof), dofs)]).max() == 0.0
