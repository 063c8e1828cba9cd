"""GPU parity of the preCICE boundary data path (/root/reference/src/boundary_condition.cpp:77-206): the nodal
coefficient field of PreciceBC (read data), the wall heat flux -k(T) grad T . n / |n| of PreciceIsothermalBC::RetrieveWriteData
(device kernel), the nodal heat-flux field in the Neumann form, against the CPU oracle.  The preCICE library and the adapter
(jots_precice.cpp) are not part of this: the tests play the adapter by writing / reading the arrays between steps."""
import numpy as np
import pytest

from cases import ini_text, write_brick_mesh, write_quad_mesh
from test_gpu_parity import APPLY_TOL, FIELD_TOL, HF0, PROPS, rel, relb, rnd, smooth_state

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def J():
    import jots_b200
    if jots_b200.load_library().jots_device_count() < 1:
        pytest.fail("no CUDA device")
    return jots_b200


def make(J, tmp_path, sim, dim, p, props, bcs, steps=3, solver="FGMRES", prec="Chebyshev", grade=0.4):
    from oracle import jots as OJ
    mesh = str(tmp_path / "m.mesh")
    if dim == 3:
        write_brick_mesh(mesh, 4, 3, 2, 0.01, 0.02, 0.015, grade=grade)
    else:
        write_quad_mesh(mesh, 6, 4, 0.02, 0.01, grade=grade)
    ini = str(tmp_path / "c.ini")
    with open(ini, "w") as f:
        f.write(ini_text(sim, mesh, 300, PROPS[props], bcs, newton=None if sim == "Linearized_Unsteady" else 50, solver=solver, prec=prec,
                         steps=steps, order=p))
    dev = J.Driver(ini)
    dev.ctx.dim = dim
    orc = OJ.JOTSDriver(OJ.Config(ini))
    assert dev.n == orc.space.ndof
    return dev, orc


def bcs_for(dim, kind, attr_index):
    nb = 6 if dim == 3 else 4
    b = ["HeatFlux, 2e5", "Isothermal, 350"] + [HF0] * (nb - 2)
    b[attr_index] = "%s, Coupling-Mesh, 320" % kind
    return b


@pytest.mark.parametrize("dim,p", [(3, 1), (3, 2), (3, 3), (2, 2), (2, 4)])
@pytest.mark.parametrize("attr_index", [0, 3])
def test_wall_heat_flux_and_node_lists(J, tmp_path, dim, p, attr_index):
    """RetrieveWriteData of the isothermal kind: same node list (DOFs, coordinates, duplicates kept) and the same wall heat
    flux as the oracle on a graded mesh, k = k(T)."""
    dev, orc = make(J, tmp_path, "Nonlinear_Unsteady", dim, p, "k2", bcs_for(dim, "preCICE_Isothermal", attr_index))
    dofs, xyz = dev.ctx.precice_bc_nodes(attr_index)
    odofs, oxyz = orc.space.bdr_nodes(attr_index + 1)
    assert (dofs == odofs).all() and np.abs(xyz - oxyz).max() < 1e-15
    u = smooth_state(orc)
    q = dev.ctx.precice_bc_write_data(attr_index, u)
    q_ref = orc.bcs[attr_index].write_data(orc.space, u, orc.props["k"])
    assert q.shape == q_ref.shape and np.abs(q_ref).max() > 0
    assert rel(q, q_ref) < APPLY_TOL, rel(q, q_ref)
    # the heat-flux kind writes the boundary temperatures
    dev2, orc2 = make(J, tmp_path, "Nonlinear_Unsteady", dim, p, "k2", bcs_for(dim, "preCICE_HeatFlux", attr_index))
    assert (dev2.ctx.precice_bc_write_data(attr_index, u) == u[odofs]).all()


@pytest.mark.parametrize("sim,props", [("Linearized_Unsteady", "k"), ("Nonlinear_Unsteady", "k"), ("Nonlinear_Unsteady", "full")])
@pytest.mark.parametrize("dim,p", [(3, 2), (2, 3)])
def test_nodal_heat_flux_in_the_neumann_forms(J, tmp_path, sim, props, dim, p):
    """PreciceHeatFluxBC: the flux is a nodal field (GridFunctionCoefficient) -- Neumann vector, and with rho(T), C(T) the
    boundary part of the residual and of the Jacobian (g / (rho_h C_h) and its derivative as ratios of interpolants)."""
    a = 0
    dev, orc = make(J, tmp_path, sim, dim, p, props, bcs_for(dim, "preCICE_HeatFlux", a))
    dofs, xyz = dev.ctx.precice_bc_nodes(a)
    g = -3e5 * (1.0 + 0.5 * np.sin(300.0 * xyz[:, 1]) * np.cos(200.0 * xyz[:, -1]))     # smooth, node-consistent read data
    dev.ctx.precice_bc_set_read_data(a, g)
    orc.bcs[a].set_read_data(orc.space, g)
    dev.op.update_neumann()
    orc.it.update_neumann()
    assert rel(dev.op.neumann_vector(), orc.it.b_vec) < APPLY_TOL
    if sim == "Nonlinear_Unsteady":
        n = dev.n
        un = smooth_state(orc)
        k = 50.0 * rnd(n, 2)
        x = rnd(n, 3)
        it = orc.it
        it.u_n = un
        it.new_state(k)
        r = dev.op.form_apply(J.FORM_RESIDUAL, k, np.empty(n), state=un)
        assert rel(r, it.mult(k)) < APPLY_TOL
        Jm = it.gradient(k)
        y = dev.op.form_apply(J.FORM_SYSTEM, x, np.empty(n), state=un + it.dt * k)
        assert relb(it.ess, y, Jm @ x) < APPLY_TOL
        d = dev.op.form_diagonal(J.FORM_SYSTEM, np.empty(n), state=un + it.dt * k)
        assert relb(it.ess, d, Jm.diagonal()) < APPLY_TOL


@pytest.mark.parametrize("kind", ["preCICE_Isothermal", "preCICE_HeatFlux"])
def test_coupled_like_time_steps(J, tmp_path, kind):
    """The adapter's loop played by the test: new read data before every step (a function of position and time), write data
    after it; fields, wall heat flux and iteration counts against the oracle doing the same."""
    a = 3
    dev, orc = make(J, tmp_path, "Nonlinear_Unsteady", 3, 2, "kc", bcs_for(3, kind, a), steps=10 ** 6)
    dofs, xyz = dev.ctx.precice_bc_nodes(a)
    for step in range(3):
        t = 0.01 * step
        if kind == "preCICE_Isothermal":
            data = 320.0 + 40.0 * np.sin(200.0 * xyz[:, 1] + 30.0 * t) * np.cos(150.0 * xyz[:, 2])
        else:
            data = -2e5 * (1.0 + 0.3 * np.sin(200.0 * xyz[:, 1] + 30.0 * t))
        dev.ctx.precice_bc_set_read_data(a, data)
        orc.bcs[a].set_read_data(orc.space, data)
        assert dev.run(1) == 1
        orc.step()
        w = dev.ctx.precice_bc_write_data(a, dev.get_u())
        w_ref = orc.bcs[a].write_data(orc.space, orc.u, orc.props["k"])
        assert rel(w, w_ref) < 1e-7
    assert rel(dev.get_u(), orc.u) < FIELD_TOL
    assert dev.ctx.stats()["newton_iters"] == orc.it.stats["newton_iters"]


@pytest.mark.parametrize("dim,p", [(3, 2), (2, 3), (3, 3)])
def test_wall_heat_flux_on_caller_supplied_tables(J, tmp_path, dim, p):
    """A space built from the caller's tables (jots_space_create_from_tables: no mesh, no boundary-element adjacency, boundary
    faces listing their DOFs in an order of their own) finds the adjacent element of every boundary face through the DOF ids
    and returns the wall heat flux at the nodes in the caller's order."""
    dev, orc = make(J, tmp_path, "Nonlinear_Unsteady", dim, p, "k2", bcs_for(dim, "preCICE_Isothermal", 0))
    s = dev.space()
    g = s.gather()
    ec, bc = s.corners()
    bg, ba = s.boundary()[:2]
    D = p + 1
    nfl = D ** (dim - 1)
    # every boundary face lists its nodes backwards (and, in 3-D, transposed): a rotation / reflection of the face lattice
    if dim == 3:
        perm = np.arange(nfl).reshape(D, D).T[::-1, ::-1].ravel()
        cperm = np.arange(4).reshape(2, 2).T[::-1, ::-1].ravel()
    else:
        perm = np.arange(nfl)[::-1]
        cperm = np.arange(2)[::-1]
    bg2 = np.ascontiguousarray(bg.reshape(-1, nfl)[:, perm])
    bc2 = np.ascontiguousarray(bc.reshape(bg2.shape[0], 2 ** (dim - 1), dim)[:, cperm, :])
    s2 = J.Space.from_tables(dim, p, s.n_dof, g, ec, bg2, bc2, ba)
    ctx = J.Context(s2)
    ctx.dim = dim
    ctx.set_material(J.THERMAL_CONDUCTIVITY, ("Polynomial", 1e-4, 0.09, -17.0))      # PROPS["k2"]
    nb = 6 if dim == 3 else 4
    ctx.set_bcs([("preCICE_Isothermal", "Coupling-Mesh", 320.0)] + [("HeatFlux", 0.0)] * (nb - 1))
    u = smooth_state(orc)
    q2 = ctx.precice_bc_write_data(0, u)
    q_ref = orc.bcs[0].write_data(orc.space, u, orc.props["k"])
    assert rel(q2.reshape(-1, nfl), q_ref.reshape(-1, nfl)[:, perm]) < APPLY_TOL
    dofs2, _ = ctx.precice_bc_nodes(0)
    dofs, _ = dev.ctx.precice_bc_nodes(0)
    assert (dofs2.reshape(-1, nfl) == dofs.reshape(-1, nfl)[:, perm]).all()
