"""GPU parity of the row-structured (TMA bulk-copy) element kernel, apply_slab4.cuh: in-memory bricks (lexicographic
numbering) with JOTS_SLAB_VARIANT=5, i.e. the kernel is forced even when the rows are shorter than a tile.  Same
checks and tolerances as tests/test_gpu_parity.py::test_axis_aligned_fast_path; every test also asserts that the
row-structured kernel really ran (stats()['rowtile_applies'])."""
import numpy as np
import pytest

from cases import ini_text
from test_gpu_parity import APPLY_TOL, FIELD_TOL, HF0, PROPS, rel, rnd, smooth_state

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def J():
    import jots_b200
    if jots_b200.load_library().jots_device_count() < 1:
        pytest.fail("no CUDA device: the GPU parity tests cannot run (there is no CPU fallback)")
    return jots_b200


def _grade(t, g):
    return t + g * t * (1.0 - t)


def build_brick(J, tmp_path, monkeypatch, sim, p, props, n, graded, variant="5", bcs=None, solver="FGMRES", prec="Chebyshev",
                steps=2, elems=None):
    from oracle import jots as OJ, mesh as OM
    monkeypatch.setenv("JOTS_SLAB_VARIANT", variant)
    if elems:
        monkeypatch.setenv("JOTS_SLAB4_ELEMS", str(elems))
    L = (0.02, 0.01, 0.015)
    dm = J.Mesh.brick(n[0], n[1], n[2], *L)
    om = OM.make_brick(n[0], n[1], n[2], *L)
    if graded:
        v = dm.arrays()[0]
        assert np.allclose(v, om.verts)
        for d, g in enumerate((0.6, -0.5, 0.3)):
            v[:, d] = L[d] * _grade(v[:, d] / L[d], g)
        dm.set_vertices(v)
        om.verts[:] = v
    if bcs is None:
        bcs = ["HeatFlux, 7.5e5", "Isothermal, 350", HF0, HF0, HF0, "HeatFlux, -2e5"]
    newton = None if sim == "Linearized_Unsteady" else 50
    txt = ini_text(sim, "unused.mesh", 300, PROPS[props], bcs, newton=newton, solver=solver, prec=prec, steps=steps, dt="1e-2", order=p)
    ini = str(tmp_path / "c.ini")
    with open(ini, "w") as f:
        f.write(txt)
    dev = J.Driver(ini, 0, mesh=dm)
    orc = OJ.JOTSDriver(OJ.Config(ini), mesh=om)
    assert dev.n == orc.space.ndof
    return dev, orc


# rows of 37 elements: a full tile plus a ragged one for every order (E = 32 / 16 / 12)
@pytest.mark.parametrize("graded", [False, True])
@pytest.mark.parametrize("p", [1, 2, 3])
@pytest.mark.parametrize("props", ["const", "k", "kc", "full"])
@pytest.mark.parametrize("sim", ["Linearized_Unsteady", "Nonlinear_Unsteady"])
def test_row_structured_kernel_forms(J, tmp_path, monkeypatch, sim, props, p, graded):
    if p == 3 and props in ("kc", "full") and sim == "Nonlinear_Unsteady":
        pytest.skip("rho(T) / C(T) in the nonlinear operator: Q1-Q2 only on the slab kernels")
    dev, orc = build_brick(J, tmp_path, monkeypatch, sim, p, props, (37, 2, 2), graded)
    n = dev.n
    state = smooth_state(orc)
    x = rnd(n, 11)
    it = orc.it
    dev.ctx.reset_stats()
    if sim == "Linearized_Unsteady":
        orc.u = state.copy()
        orc.update_mat_props(True)
        for form, mat in ((J.FORM_MASS, it.M_mat), (J.FORM_STIFFNESS, it.K_mat), (J.FORM_SYSTEM, it.A_mat)):
            y = dev.op.form_apply(form, x, np.empty(n), state=state)
            assert rel(y, mat @ x) < APPLY_TOL, (form, rel(y, mat @ x))
        assert dev.ctx.stats()["rowtile_applies"] == 3
    else:
        k = 50.0 * rnd(n, 12)
        it.u_n = state
        it.new_state(k)
        r = dev.op.form_apply(J.FORM_RESIDUAL, k, np.empty(n), state=state)
        assert rel(r, it.mult(k)) < APPLY_TOL
        Jm = it.gradient(k)
        y = dev.op.form_apply(J.FORM_SYSTEM, x, np.empty(n), state=state + it.dt * k)
        assert rel(y, Jm @ x) < APPLY_TOL
        assert dev.ctx.stats()["rowtile_applies"] == 2


@pytest.mark.parametrize("elems", [14, 16])
def test_row_structured_kernel_tile_lengths_q2(J, tmp_path, monkeypatch, elems):
    """Both Q2 tile lengths, rows that are an exact multiple / not a multiple of the tile, all faces essential or not."""
    for nx, bcs in ((32, None), (45, ["Isothermal, 350"] * 6), (5, ["Isothermal, 320", HF0, "Isothermal, 400", HF0, HF0, HF0])):
        dev, orc = build_brick(J, tmp_path, monkeypatch, "Nonlinear_Unsteady", 2, "k", (nx, 3, 2), True, bcs=bcs)
        n = dev.n
        state = smooth_state(orc)
        x, k = rnd(n, 21), 50.0 * rnd(n, 22)
        it = orc.it
        it.u_n = state
        it.new_state(k)
        dev.ctx.reset_stats()
        r = dev.op.form_apply(J.FORM_RESIDUAL, k, np.empty(n), state=state)
        assert rel(r, it.mult(k)) < APPLY_TOL
        y = dev.op.form_apply(J.FORM_SYSTEM, x, np.empty(n), state=state + it.dt * k)
        assert rel(y, it.gradient(k) @ x) < APPLY_TOL
        assert dev.ctx.stats()["rowtile_applies"] == 2


@pytest.mark.parametrize("sim,props,solver,prec,p", [
    ("Nonlinear_Unsteady", "k", "CG", "Jacobi", 2),
    ("Nonlinear_Unsteady", "kc", "FGMRES", "Chebyshev", 2),
    ("Linearized_Unsteady", "full", "GMRES", "Jacobi", 2),
    ("Nonlinear_Unsteady", "k", "FGMRES", "Chebyshev", 3),
    ("Linearized_Unsteady", "k", "CG", "Jacobi", 1),
])
def test_row_structured_kernel_time_steps(J, tmp_path, monkeypatch, sim, props, solver, prec, p):
    """Whole implicit steps (Krylov vectors at odd 8-byte offsets, many applies): field within 1e-8 of the oracle."""
    bcs = ["HeatFlux, 7.5e5", HF0, HF0, "Isothermal, 300", HF0, HF0]
    dev, orc = build_brick(J, tmp_path, monkeypatch, sim, p, props, (36, 3, 2), False, bcs=bcs, solver=solver, prec=prec, steps=2)
    assert dev.run() == 2
    u_ref = orc.run()
    assert rel(dev.get_u(), u_ref) < FIELD_TOL
    assert dev.ctx.stats()["rowtile_applies"] > 0


def test_row_structured_kernel_matches_gather_kernel_at_size(J, tmp_path, monkeypatch):
    """48^3 Q2 bricks (912 673 DOFs, every block of the persistent grid loops over several tiles): the row-structured
    kernel and the gather-table kernel agree to summation order on Jacobian, residual and linear applies."""
    import jots_b200
    out = {}
    for variant in ("3", "4"):
        monkeypatch.setenv("JOTS_SLAB_VARIANT", variant)
        txt = ini_text("Nonlinear_Unsteady", "unused.mesh", 300, PROPS["k"], ["HeatFlux, 7.5e5", HF0, HF0, "Isothermal, 300", HF0, HF0],
                       newton=50, solver="CG", prec="Jacobi", steps=1, dt="1e-2", order=2)
        ini = str(tmp_path / ("c%s.ini" % variant))
        with open(ini, "w") as f:
            f.write(txt)
        dev = jots_b200.Driver(ini, 0, mesh=jots_b200.Mesh.brick(48, 48, 48, 0.01, 0.01, 0.01))
        n = dev.n
        X = dev.space().node_coordinates()
        state = 300.0 + 100.0 * np.prod(np.sin(np.pi * X / 0.01), axis=1)
        x = rnd(n, 5)
        dev.ctx.reset_stats()
        out[variant] = (dev.op.form_apply(J.FORM_SYSTEM, x, np.empty(n), state=state),
                        dev.op.form_apply(J.FORM_RESIDUAL, x, np.empty(n), state=state))
        assert (dev.ctx.stats()["rowtile_applies"] > 0) == (variant == "4")
        del dev
    for a, b in zip(out["3"], out["4"]):
        assert rel(b, a) < 1e-13


def test_row_structured_kernel_unaligned_vectors(J, tmp_path, monkeypatch):
    """Device vectors that start on an odd double (8-byte aligned only): the bulk copies shift their window by one
    double, the first and last rows of the vector take the plain-load path."""
    import torch
    dev, orc = build_brick(J, tmp_path, monkeypatch, "Nonlinear_Unsteady", 2, "k", (37, 3, 2), False)
    n = dev.n
    x, state = rnd(n, 31), smooth_state(orc)
    y_ref = dev.op.form_apply(J.FORM_SYSTEM, x, np.empty(n), state=state)
    for ox, oz, oy in ((1, 0, 0), (0, 1, 1), (1, 1, 0)):
        bx = torch.zeros(n + 2, dtype=torch.float64, device="cuda")
        bz = torch.zeros(n + 2, dtype=torch.float64, device="cuda")
        by = torch.zeros(n + 2, dtype=torch.float64, device="cuda")
        xd, zd, yd = bx[ox:ox + n], bz[oz:oz + n], by[oy:oy + n]
        xd.copy_(torch.from_numpy(x))
        zd.copy_(torch.from_numpy(state))
        torch.cuda.synchronize()
        assert (xd.data_ptr() // 8) % 2 == ox and (zd.data_ptr() // 8) % 2 == oz
        dev.ctx.reset_stats()
        dev.op.form_apply(J.FORM_SYSTEM, xd, yd, state=zd)
        assert dev.ctx.stats()["rowtile_applies"] == 1
        assert rel(yd.cpu().numpy(), y_ref) < 1e-13
        assert float(by[:oy].abs().sum()) == 0.0 and float(by[oy + n:].abs().sum()) == 0.0
