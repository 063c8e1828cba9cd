"""GPU parity of the unique-pencil patch kernel (jots_b200/csrc/apply_patch.cuh, the default element kernel for Q2 hexes on
lattice meshes) against the CPU oracle, through the C ABI: full and ragged patches, holes, graded geometry (one Jacobian
record per element), essential faces, every form the kernel is instantiated for, and whole time steps.  The build helper
writes a mesh FILE, so the lattice is recovered from a first-appearance DOF numbering; the brick / partitioned paths are
covered by test_gpu_parity.py::test_fast_path_persistent_loop_against_the_oracle, smoke() and test_gpu_dist.py."""
import os

import numpy as np
import pytest

from test_gpu_parity import APPLY_TOL, FIELD_TOL, HF0, build, rel, relb, rnd, smooth_state

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def J():
    import jots_b200
    if jots_b200.load_library().jots_device_count() < 1:
        pytest.fail("no CUDA device")
    return jots_b200


@pytest.fixture(autouse=True)
def _force_patches():
    """ragged patches too: the default keeps the gather-table kernel below a fill of 0.6"""
    old = {k: os.environ.get(k) for k in ("JOTS_PATCH_MIN_FILL", "JOTS_SLAB_VARIANT")}
    os.environ["JOTS_PATCH_MIN_FILL"] = "0"
    os.environ.pop("JOTS_SLAB_VARIANT", None)
    yield
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("props", ["const", "k", "k2", "kc", "full"])
@pytest.mark.parametrize("sim", ["Linearized_Unsteady", "Nonlinear_Unsteady"])
@pytest.mark.parametrize("n", [(5, 3, 4), (8, 8, 3), (1, 1, 1), (13, 9, 2)])
def test_patch_kernel_forms(J, tmp_path, sim, props, n):
    dev, orc = build(J, tmp_path, sim, 3, 2, props, grade=0.6, shear=0.0, n=n)
    st0 = dev.ctx.stats()
    assert st0["n_patches"] == -(-n[0] // 4) * -(-n[1] // 4) * n[2]
    nd = dev.n
    state = smooth_state(orc)
    x = rnd(nd, 11)
    it = orc.it
    if sim == "Linearized_Unsteady":
        orc.u = state.copy()
        orc.update_mat_props(True)
        for form, mat in ((J.FORM_MASS, it.M_mat), (J.FORM_STIFFNESS, it.K_mat), (J.FORM_SYSTEM, it.A_mat)):
            y = dev.op.form_apply(form, x, np.empty(nd), state=state)
            assert relb(it.ess, y, mat @ x) < APPLY_TOL, (form, relb(it.ess, y, mat @ x))
        assert dev.ctx.stats()["patch_applies"] - st0["patch_applies"] == 3
    else:
        k = 50.0 * rnd(nd, 12)
        it.u_n = state
        it.new_state(k)
        r = dev.op.form_apply(J.FORM_RESIDUAL, k, np.empty(nd), state=state)
        assert rel(r, it.mult(k)) < APPLY_TOL, rel(r, it.mult(k))
        Jm = it.gradient(k)
        y = dev.op.form_apply(J.FORM_SYSTEM, x, np.empty(nd), state=state + it.dt * k)
        assert relb(it.ess, y, Jm @ x) < APPLY_TOL, relb(it.ess, y, Jm @ x)
        assert dev.ctx.stats()["patch_applies"] - st0["patch_applies"] == 2


@pytest.mark.parametrize("sim,props,solver,prec", [("Nonlinear_Unsteady", "k", "CG", "Jacobi"), ("Nonlinear_Unsteady", "full", "FGMRES", "Chebyshev"),
                                                   ("Linearized_Unsteady", "full", "GMRES", "Jacobi"), ("Steady", "k", "CG", "Chebyshev")])
def test_patch_kernel_time_steps(J, tmp_path, sim, props, solver, prec):
    """Whole steps: temperature field against the oracle, and identical Krylov / Newton counts to the gather-table kernel."""
    dev, orc = build(J, tmp_path, sim, 3, 2, props, solver=solver, prec=prec, n=(9, 6, 3), steps=3,
                     bcs=["HeatFlux, 7.5e5", "Isothermal, 350", HF0, "Sinusoidal_Isothermal, 50, 40, 0.3, 400", HF0, HF0])
    dev.run()
    st = dev.ctx.stats()
    assert st["patch_applies"] > 0
    u_ref = orc.run()
    assert rel(dev.get_u(), u_ref) < FIELD_TOL
    os.environ["JOTS_SLAB_VARIANT"] = "3"
    dev3, _ = build(J, tmp_path, sim, 3, 2, props, solver=solver, prec=prec, n=(9, 6, 3), steps=3,
                    bcs=["HeatFlux, 7.5e5", "Isothermal, 350", HF0, "Sinusoidal_Isothermal, 50, 40, 0.3, 400", HF0, HF0])
    dev3.run()
    st3 = dev3.ctx.stats()
    assert st3["patch_applies"] == 0 and st3["n_patches"] == 0
    assert st3["newton_iters"] == st["newton_iters"] and abs(st3["krylov_iters"] - st["krylov_iters"]) <= 2
    assert rel(dev3.get_u(), dev.get_u()) < 1e-10


def test_patch_kernel_on_a_brick_with_vectors_on_odd_offsets(J, tmp_path):
    """Device pointers that are only 8-byte aligned (cp.async 8-byte gathers have no stronger requirement)."""
    import torch
    from cases import ini_text
    from test_gpu_parity import PROPS
    ini = str(tmp_path / "c.ini")
    with open(ini, "w") as f:
        f.write(ini_text("Nonlinear_Unsteady", "unused", 300, PROPS["k"], ["HeatFlux, 7.5e5", HF0, HF0, "Isothermal, 300", HF0, HF0],
                         newton=50, steps=1, order=2))
    dev = J.Driver(ini, 0, mesh=J.Mesh.brick(12, 8, 5, 0.01, 0.01, 0.01))
    n = dev.n
    assert dev.ctx.stats()["n_patches"] == 3 * 2 * 5
    g = torch.Generator(device="cuda").manual_seed(5)
    buf = torch.empty(3 * (n + 1), dtype=torch.float64, device="cuda").uniform_(-1.0, 1.0, generator=g)
    x, z = buf[1:n + 1], buf[n + 2:2 * n + 2]
    z.mul_(100.0).add_(500.0)
    y0 = torch.empty(n, dtype=torch.float64, device="cuda")
    y1 = buf[2 * n + 3:3 * n + 3]
    torch.cuda.synchronize()
    dev.op.form_apply(J.FORM_SYSTEM, x.clone(), y0, state=z.clone())
    dev.op.form_apply(J.FORM_SYSTEM, x, y1, state=z)
    torch.cuda.synchronize()
    assert float((y0 - y1).abs().max()) <= 1e-12 * float(y0.abs().max())


@pytest.mark.parametrize("sim,props", [("Nonlinear_Unsteady", "k"), ("Linearized_Unsteady", "const"), ("Nonlinear_Unsteady", "const"),
                                       ("Nonlinear_Unsteady", "k2"), ("Steady", "k")])
@pytest.mark.parametrize("n", [(8, 8, 3), (5, 3, 4), (13, 9, 2)])
def test_inner_product_formed_inside_the_element_kernel(J, tmp_path, sim, props, n):
    """(x, A x) accumulated by phase 3 of the two-phase patch kernel (FormArgs::dot: pencil outputs times the staged inputs,
    unit-diagonal rows added by the zeroing kernel) equals x . (A x) of the oracle's matrix -- x with NON-zero essential entries,
    ragged patches, graded geometry; forms the two-phase kernel does not cover (quadratic k) report fused = False and the
    separate pass gives the same number."""
    dev, orc = build(J, tmp_path, sim, 3, 2, props, grade=0.6, shear=0.0, n=n,
                     bcs=["HeatFlux, 7.5e5", "Isothermal, 350", HF0, "Isothermal, 300", HF0, HF0])
    nd = dev.n
    state = smooth_state(orc)
    x = rnd(nd, 21)
    it = orc.it
    if sim == "Linearized_Unsteady":
        orc.u = state.copy()
        orc.update_mat_props(True)
        A, zs = it.A_mat, state
    elif sim == "Nonlinear_Unsteady":
        k = 50.0 * rnd(nd, 22)
        it.u_n = state
        it.new_state(k)
        A, zs = it.gradient(k), state + it.dt * k
    else:
        it.new_state(state)
        A, zs = it.gradient(state), state
    y = np.empty(nd)
    dot, fused = dev.op.form_apply_dot(J.FORM_SYSTEM, x, y, state=zs)
    assert fused == (props != "k2")
    ref = A @ x
    assert relb(it.ess, y, ref) < APPLY_TOL
    assert abs(dot - x @ ref) <= 1e-12 * np.abs(x * ref).sum(), (dot, x @ ref)
    assert dev.ctx.stats()["fused_dots"] == (1 if fused else 0)


@pytest.mark.parametrize("sim,props,n,max_it", [("Nonlinear_Unsteady", "k", (8, 8, 4), 100), ("Linearized_Unsteady", "const", (9, 5, 3), 100),
                                                ("Nonlinear_Unsteady", "k", (8, 8, 4), 7), ("Steady", "k", (8, 4, 2), 100)])
def test_cg_with_the_fused_inner_product_matches_the_separate_pass(J, tmp_path, monkeypatch, sim, props, n, max_it):
    """cg_device_loop with den = (d, A d) from the element kernel against JOTS_CG_FUSED_DOT=0: same Krylov / Newton counts,
    same field, and the oracle's field."""
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("JOTS_CG_FUSED_DOT", mode)
        dev, orc = build(J, tmp_path, sim, 3, 2, props, solver="CG", prec="Jacobi", n=n, steps=2)
        if max_it != 100:
            txt = open(str(tmp_path / "c.ini")).read().replace("Max_Iterations=100\n", "Max_Iterations=%d\n" % max_it, 1)
            open(str(tmp_path / "c.ini"), "w").write(txt)
            dev = J.Driver(str(tmp_path / "c.ini"))
        try:
            dev.run()
        except J.JotsError:
            pass
        st = dev.ctx.stats()
        assert (st["fused_dots"] > 0) == (mode == "1"), st
        out[mode] = (st["krylov_iters"], st["newton_iters"], st["last_linear_iters"], st["last_linear_converged"], dev.get_u())
    assert out["0"][:4] == out["1"][:4], (out["0"][:4], out["1"][:4])
    assert rel(out["0"][4], out["1"][4]) < 1e-12
    if max_it == 100:
        assert rel(out["0"][4], orc.run()) < FIELD_TOL


@pytest.mark.parametrize("sim,props", [("Nonlinear_Unsteady", "k"), ("Linearized_Unsteady", "const")])
@pytest.mark.parametrize("p,n", [(2, (8, 8, 3)), (2, (5, 3, 4)), (2, (13, 9, 2)), (2, (4, 4, 1)), (3, (5, 3, 4)), (3, (7, 5, 3)), (1, (13, 9, 2))])
def test_side_stream_carried_by_the_element_kernel(J, tmp_path, sim, props, p, n):
    """The apply the conjugate-gradient loop launches (FormArgs::side_*): besides y = A x and (x, A x) the block that processes
    tile t does x_sol += alpha d_old and zeroes the next output buffer over slice t of the vectors.  y and the inner product
    against the oracle's matrix, x_sol against alpha * d_old + x_sol entry by entry (one fused multiply-add each), the zeroed
    buffer NaN-filled beforehand so that a slice nobody visited shows; odd vector lengths, ragged patches.  A single patch holds
    more entries than one tile can carry: the kernel reports that it did not, and the loop keeps the vector kernels.  Q2: the
    two-phase patch kernel; Q1 and Q3: the gather-table kernel (two staging buffers)."""
    import torch
    dev, orc = build(J, tmp_path, sim, 3, p, props, grade=0.6, shear=0.0, n=n,
                     bcs=["HeatFlux, 7.5e5", "Isothermal, 350", HF0, "Isothermal, 300", HF0, HF0])
    nd = dev.n
    state = smooth_state(orc)
    x = rnd(nd, 31)
    it = orc.it
    if sim == "Linearized_Unsteady":
        orc.u = state.copy()
        orc.update_mat_props(True)
        A, zs = it.A_mat, state
    else:
        k = 50.0 * rnd(nd, 32)
        it.u_n = state
        it.new_state(k)
        A, zs = it.gradient(k), state + it.dt * k
    alpha = -0.37
    d_old, x_sol = rnd(nd, 33), 10.0 * rnd(nd, 34)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    tx, tz, td, txs = t(x), t(zs), t(d_old), t(x_sol)
    ty = torch.full((nd,), float("nan"), dtype=torch.float64, device="cuda")
    tzero = torch.full((nd,), float("nan"), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    dot, carried = dev.op.form_apply_side(J.FORM_SYSTEM, tx, ty, tz, alpha, td, txs, tzero)
    torch.cuda.synchronize()
    if n == (4, 4, 1):
        assert not carried and dev.ctx.stats()["side_applies"] == 0
        return
    assert carried and dev.ctx.stats()["side_applies"] == 1 and dev.ctx.last_kernel().endswith(",side>")
    ref = A @ x
    assert relb(it.ess, ty.cpu().numpy(), ref) < APPLY_TOL
    assert abs(dot - x @ ref) <= 1e-12 * np.abs(x * ref).sum(), (dot, x @ ref)
    got = txs.cpu().numpy()
    want = x_sol + alpha * d_old
    assert np.all(np.abs(got - want) <= 2.3e-16 * (np.abs(x_sol) + np.abs(alpha * d_old))), np.abs(got - want).max()
    assert np.array_equal(tzero.cpu().numpy(), np.zeros(nd))
    assert np.array_equal(td.cpu().numpy(), d_old) and np.array_equal(tx.cpu().numpy(), x)


@pytest.mark.parametrize("sim,props,p,n,max_it", [("Nonlinear_Unsteady", "k", 2, (8, 8, 4), 100), ("Linearized_Unsteady", "const", 2, (9, 5, 3), 100),
                                                  ("Nonlinear_Unsteady", "k", 2, (8, 8, 4), 7), ("Nonlinear_Unsteady", "k", 2, (8, 8, 4), 1),
                                                  ("Steady", "k", 2, (8, 4, 2), 100), ("Nonlinear_Unsteady", "k", 3, (5, 3, 4), 100),
                                                  ("Nonlinear_Unsteady", "k", 3, (5, 3, 4), 6), ("Linearized_Unsteady", "k", 1, (13, 9, 2), 100)])
def test_cg_with_the_side_stream_matches_the_plain_loop(J, tmp_path, monkeypatch, sim, props, p, n, max_it):
    """cg_device_loop with x += alpha d and the zeroing of the next output carried by the element kernel against JOTS_CG_SIDE=0
    (both in the vector kernels): every vector entry sees the same operations in the same order (only the summation order of the
    element kernel's atomics varies from launch to launch), so the counts are equal and norms and field agree to rounding --
    early convergence, a stop inside a chunk, Max_Iterations (odd and even, so both buffers of the ping-pong end up holding the
    pending update) -- and the field is the oracle's."""
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("JOTS_CG_SIDE", mode)
        dev, orc = build(J, tmp_path, sim, 3, p, props, solver="CG", prec="Jacobi", n=n, steps=2)
        if max_it != 100:
            txt = open(str(tmp_path / "c.ini")).read().replace("Max_Iterations=100\n", "Max_Iterations=%d\n" % max_it, 1)
            open(str(tmp_path / "c.ini"), "w").write(txt)
            dev = J.Driver(str(tmp_path / "c.ini"))
        try:
            dev.run()
        except J.JotsError:
            pass
        st = dev.ctx.stats()
        assert (st["side_applies"] > 0) == (mode == "1" and max_it > 1), st
        assert st["fused_dots"] > 0 or max_it == 1
        out[mode] = (st["krylov_iters"], st["newton_iters"], st["last_linear_iters"], st["last_linear_converged"], st["last_linear_norm"],
                     dev.get_u())
    assert out["0"][:4] == out["1"][:4], (out["0"][:4], out["1"][:4])
    # the final (B r, r) sits 20 orders of magnitude below the first one: the atomics' summation order shows in its digits
    assert abs(out["0"][4] - out["1"][4]) <= 1e-4 * abs(out["0"][4])
    assert rel(out["0"][5], out["1"][5]) < 1e-12
    if max_it == 100:
        assert rel(out["0"][5], orc.run()) < FIELD_TOL
