"""GPU parity: every device form / solver / driver path against the CPU oracle on the same inputs,
called through the C ABI (ctypes).  Tolerances: FP64, relative l2 <= 1e-11 for single operator
applications (summation-order differences only), <= 1e-8 for converged temperature fields (the
north_star bound, same solver rel_tol on both sides)."""
import numpy as np
import pytest

from cases import CASES, ini_text, make_case, write_brick_mesh, write_quad_mesh

pytestmark = pytest.mark.gpu

APPLY_TOL = 1e-11
FIELD_TOL = 1e-8


@pytest.fixture(scope="module")
def J():
    import jots_b200
    lib = jots_b200.load_library()
    if lib.jots_device_count() < 1:
        pytest.fail("no CUDA device: the GPU parity tests cannot run (there is no CPU fallback)")
    return jots_b200


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def relb(ess, a, b):
    """relative l2 error taken SEPARATELY on the essential rows (unit diagonal: entries O(1)) and on the free rows (entries
    O(h^3), eight orders of magnitude smaller on these meshes), so that the large block cannot hide an error in the small one"""
    ess = np.asarray(ess, dtype=np.int64)
    free = np.ones(len(b), dtype=bool)
    free[ess] = False
    e = rel(a[free], b[free]) if free.any() else 0.0
    return max(e, rel(a[ess], b[ess])) if ess.size else e


def rnd(n, seed):
    return np.random.default_rng(seed).uniform(-1.0, 1.0, n)


HF0 = "HeatFlux, 0"
PROPS = {
    "const": [("Density_rho", "Uniform, 8000"), ("Specific_Heat_C", "Uniform, 500"), ("Thermal_Conductivity_k", "Uniform, 10")],
    "k": [("Density_rho", "Uniform, 8000"), ("Specific_Heat_C", "Uniform, 500"), ("Thermal_Conductivity_k", "Polynomial, 0.09, -17")],
    "k2": [("Density_rho", "Uniform, 8000"), ("Specific_Heat_C", "Uniform, 500"), ("Thermal_Conductivity_k", "Polynomial, 1e-4, 0.09, -17")],
    "kc": [("Density_rho", "Uniform, 8000"), ("Specific_Heat_C", "Polynomial, 4.5, -850"),
           ("Thermal_Conductivity_k", "Polynomial, 1e-4, 0.09, -17")],
    "full": [("Density_rho", "Polynomial, 1e-3, -0.5, 8000"), ("Specific_Heat_C", "Polynomial, 4.5, -850"),
             ("Thermal_Conductivity_k", "Polynomial, 1e-4, 0.09, -17")],
}


def build(J, tmp_path, sim, dim, p, props, bcs=None, grade=0.0, shear=0.0, solver="FGMRES", prec="Chebyshev", n=(3, 2, 2),
          newton=None, scheme="Euler_Implicit", dt="1e-2", t0=300, steps=3, jitter=0.0):
    """Write mesh + ini, build the device driver and the oracle driver on them."""
    from oracle import jots as OJ
    mesh = str(tmp_path / "m.mesh")
    if dim == 3:
        write_brick_mesh(mesh, n[0], n[1], n[2], 0.01, 0.02, 0.015, grade=grade, shear=shear, jitter=jitter)
        nb = 6
    else:
        write_quad_mesh(mesh, n[0] + 1, n[1] + 1, 0.02, 0.01, grade=grade, shear=shear, jitter=jitter)
        nb = 4
    if bcs is None:
        bcs = ["HeatFlux, 7.5e5", "Isothermal, 350"] + [HF0] * (nb - 3) + ["HeatFlux, -2e5"]
    if newton is None and sim != "Linearized_Unsteady":
        newton = 50
    txt = ini_text(sim, mesh, t0, PROPS[props] if sim != "Steady" else [PROPS[props][2]], bcs, time=(sim != "Steady"),
                   newton=newton, solver=solver, prec=prec, steps=steps, dt=dt, order=p, scheme=scheme)
    ini = str(tmp_path / "c.ini")
    with open(ini, "w") as f:
        f.write(txt)
    dev = J.Driver(ini)
    orc = OJ.JOTSDriver(OJ.Config(ini))
    assert dev.n == orc.space.ndof
    return dev, orc


def smooth_state(orc, lo=320.0, hi=900.0):
    X = orc.space.node_coordinates()
    s = np.ones(X.shape[0])
    for d in range(X.shape[1]):
        L = X[:, d].max() - X[:, d].min()
        s *= 0.5 + 0.5 * np.sin(2.3 * (X[:, d] - X[:, d].min()) / L + 0.4 * d)
    return lo + (hi - lo) * s


SHAPES = [(3, 1), (3, 2), (3, 3), (3, 4), (2, 1), (2, 2), (2, 3), (2, 4)]


@pytest.mark.parametrize("dim,p", SHAPES)
@pytest.mark.parametrize("props", ["const", "k", "full"])
def test_linearized_forms(J, tmp_path, dim, p, props):
    """M, K, A = M + dt K of LinearConductionOperator (linear_conduction_operator.cpp:59-120)."""
    dev, orc = build(J, tmp_path, "Linearized_Unsteady", dim, p, props, grade=0.6, shear=0.3)
    n = dev.n
    state = smooth_state(orc)
    x = rnd(n, 1)
    # oracle: re-assemble at the state
    orc.u = state.copy()
    orc.update_mat_props(True)
    it = orc.it
    for form, mat in ((J.FORM_MASS, it.M_mat), (J.FORM_STIFFNESS, it.K_mat), (J.FORM_SYSTEM, it.A_mat)):
        y = dev.op.form_apply(form, x, np.empty(n), state=state)
        assert relb(it.ess, y, mat @ x) < APPLY_TOL, (form, relb(it.ess, y, mat @ x))
        if form != J.FORM_STIFFNESS:
            d = dev.op.form_diagonal(form, np.empty(n), state=state)
            assert relb(it.ess, d, mat.diagonal()) < APPLY_TOL
    assert (dev.ctx.essential_dofs() == it.ess).all()
    assert rel(dev.op.neumann_vector(), it.b_vec) < APPLY_TOL


@pytest.mark.parametrize("dim,p", SHAPES)
@pytest.mark.parametrize("props", ["const", "k", "full"])
def test_nonlinear_forms(J, tmp_path, dim, p, props):
    """R(k) and its Jacobian (nl_conduction_operator.cpp:27-90; jots_nlfis.cpp:4-95)."""
    dev, orc = build(J, tmp_path, "Nonlinear_Unsteady", dim, p, props, grade=0.6, shear=0.3)
    n = dev.n
    un = smooth_state(orc)
    k = 50.0 * rnd(n, 2)
    x = rnd(n, 3)
    it = orc.it
    it.u_n = un
    it.new_state(k)
    r_ref = it.mult(k)
    Jm = it.gradient(k)
    z = un + it.dt * k
    r = dev.op.form_apply(J.FORM_RESIDUAL, k, np.empty(n), state=un)
    assert rel(r, r_ref) < APPLY_TOL, rel(r, r_ref)
    y = dev.op.form_apply(J.FORM_SYSTEM, x, np.empty(n), state=z)
    assert relb(it.ess, y, Jm @ x) < APPLY_TOL, relb(it.ess, y, Jm @ x)
    d = dev.op.form_diagonal(J.FORM_SYSTEM, np.empty(n), state=z)
    assert relb(it.ess, d, Jm.diagonal()) < APPLY_TOL
    y = dev.op.form_apply(J.FORM_MASS, x, np.empty(n))
    assert relb(it.ess, y, it.M_mat @ x) < APPLY_TOL
    assert rel(dev.op.neumann_vector(), it.b_vec) < APPLY_TOL


@pytest.mark.parametrize("p", [1, 2, 3])
@pytest.mark.parametrize("props", ["const", "k", "k2", "kc", "full"])
@pytest.mark.parametrize("sim", ["Linearized_Unsteady", "Nonlinear_Unsteady"])
def test_axis_aligned_fast_path(J, tmp_path, sim, props, p):
    """Axis-aligned (graded) bricks take the register-slab kernel: same checks as above.  "k" (k linear in T) runs the
    Jacobian kernel with a uniform k' (CF_KL), "k2" (quadratic k) the general one."""
    dev, orc = build(J, tmp_path, sim, 3, p, props, grade=0.6, shear=0.0, n=(5, 3, 4))
    n = dev.n
    state = smooth_state(orc)
    x = rnd(n, 11)
    it = orc.it
    if sim == "Linearized_Unsteady":
        orc.u = state.copy()
        orc.update_mat_props(True)
        for form, mat in ((J.FORM_MASS, it.M_mat), (J.FORM_STIFFNESS, it.K_mat), (J.FORM_SYSTEM, it.A_mat)):
            y = dev.op.form_apply(form, x, np.empty(n), state=state)
            assert relb(it.ess, y, mat @ x) < APPLY_TOL, (form, relb(it.ess, y, mat @ x))
            if form != J.FORM_STIFFNESS:
                d = dev.op.form_diagonal(form, np.empty(n), state=state)
                assert relb(it.ess, d, mat.diagonal()) < APPLY_TOL, (form, relb(it.ess, d, mat.diagonal()))
    else:
        k = 50.0 * rnd(n, 12)
        it.u_n = state
        it.new_state(k)
        r = dev.op.form_apply(J.FORM_RESIDUAL, k, np.empty(n), state=state)
        assert rel(r, it.mult(k)) < APPLY_TOL
        Jm = it.gradient(k)
        y = dev.op.form_apply(J.FORM_SYSTEM, x, np.empty(n), state=state + it.dt * k)
        assert relb(it.ess, y, Jm @ x) < APPLY_TOL
        d = dev.op.form_diagonal(J.FORM_SYSTEM, np.empty(n), state=state + it.dt * k)
        assert relb(it.ess, d, Jm.diagonal()) < APPLY_TOL, relb(it.ess, d, Jm.diagonal())


@pytest.mark.parametrize("p,n,props", [(2, (24, 24, 20), "k"), (2, (24, 24, 20), "full"), (3, (16, 16, 15), "k"), (1, (40, 40, 32), "k2")])
def test_fast_path_persistent_loop_against_the_oracle(J, tmp_path, p, n, props):
    """More tiles than resident blocks (Q2: 11 520 elements = 720 16-element tiles on at most 592 blocks): every block of
    the persistent kernel takes the `it >= 1` path -- refill of the single staging buffer, rotation of the index buffers --
    and the result is compared with the ORACLE's assembled sparse Jacobian / residual / diagonal, not with another kernel."""
    dev, orc = build(J, tmp_path, "Nonlinear_Unsteady", 3, p, props, grade=0.6, shear=0.0, n=n, steps=1)
    nd = dev.n
    state = smooth_state(orc)
    x = rnd(nd, 31)
    k = 50.0 * rnd(nd, 32)
    it = orc.it
    it.u_n = state
    it.new_state(k)
    r = dev.op.form_apply(J.FORM_RESIDUAL, k, np.empty(nd), state=state)
    assert rel(r, it.mult(k)) < APPLY_TOL
    Jm = it.gradient(k)
    z = state + it.dt * k
    y = dev.op.form_apply(J.FORM_SYSTEM, x, np.empty(nd), state=z)
    assert relb(it.ess, y, Jm @ x) < APPLY_TOL, relb(it.ess, y, Jm @ x)
    d = dev.op.form_diagonal(J.FORM_SYSTEM, np.empty(nd), state=z)
    assert relb(it.ess, d, Jm.diagonal()) < APPLY_TOL, relb(it.ess, d, Jm.diagonal())


@pytest.mark.parametrize("dim,p,shear", [(3, 2, 0.2), (3, 2, 0.0), (3, 1, 0.0), (3, 3, 0.2), (2, 2, 0.2), (2, 1, 0.2)])
def test_steady_forms(J, tmp_path, dim, p, shear):
    dev, orc = build(J, tmp_path, "Steady", dim, p, "k", grade=0.4, shear=shear, solver="CG")
    n = dev.n
    u = smooth_state(orc)
    x = rnd(n, 4)
    it = orc.it
    it.new_state(u)
    r = dev.op.form_apply(J.FORM_RESIDUAL, u, np.empty(n))
    assert rel(r, it.mult(u)) < APPLY_TOL
    Jm = it.gradient(u)
    y = dev.op.form_apply(J.FORM_SYSTEM, x, np.empty(n), state=u)
    assert relb(it.ess, y, Jm @ x) < APPLY_TOL
    d = dev.op.form_diagonal(J.FORM_SYSTEM, np.empty(n), state=u)
    assert relb(it.ess, d, Jm.diagonal()) < APPLY_TOL


@pytest.mark.parametrize("solver", ["CG", "GMRES", "FGMRES"])
@pytest.mark.parametrize("prec", ["Jacobi", "Chebyshev"])
def test_linearized_steps_all_solvers(J, tmp_path, solver, prec):
    """Three implicit steps with k(T), C(T): fields and iteration counts against the oracle."""
    dev, orc = build(J, tmp_path, "Linearized_Unsteady", 3, 2, "full", solver=solver, prec=prec, n=(6, 3, 2), steps=3,
                     bcs=["HeatFlux, 7.5e5", "Isothermal, 350", HF0, "Sinusoidal_Isothermal, 50, 40, 0.3, 400", HF0, HF0])
    assert dev.run() == 3
    u_ref = orc.run()
    assert rel(dev.get_u(), u_ref) < FIELD_TOL
    st = dev.ctx.stats()
    assert st["steps"] == 3 and st["last_linear_converged"] == 1
    assert abs(st["krylov_iters"] - orc.it.stats["krylov_iters"]) <= 3


@pytest.mark.parametrize("solver,prec", [("FGMRES", "Chebyshev"), ("GMRES", "Jacobi"), ("CG", "Jacobi")])
@pytest.mark.parametrize("props", ["const", "k", "full"])
def test_nonlinear_steps(J, tmp_path, solver, prec, props):
    # CG on the non-symmetric Jacobian of the k(T) / C(T) cases is what BASELINE configs[3] ("Newton-CG") and bench.py
    # run: MFEM's CG control flow is followed on both sides, so fields and Newton counts must still agree
    dev, orc = build(J, tmp_path, "Nonlinear_Unsteady", 3, 2, props, solver=solver, prec=prec, n=(6, 3, 2), steps=3)
    assert dev.run() == 3
    u_ref = orc.run()
    assert rel(dev.get_u(), u_ref) < FIELD_TOL
    st = dev.ctx.stats()
    assert st["last_newton_converged"] == 1
    assert st["newton_iters"] == orc.it.stats["newton_iters"]


@pytest.mark.parametrize("scheme", ["Euler_Explicit", "RK4"])
@pytest.mark.parametrize("sim", ["Linearized_Unsteady", "Nonlinear_Unsteady"])
def test_explicit_paths(J, tmp_path, sim, scheme):
    dev, orc = build(J, tmp_path, sim, 2, 2, "const", scheme=scheme, dt="1e-5", steps=4, n=(4, 3, 1))
    assert dev.run() == 4
    assert rel(dev.get_u(), orc.run()) < FIELD_TOL


@pytest.mark.parametrize("scheme", ["Euler_Explicit", "RK4"])
@pytest.mark.parametrize("dim,props", [(3, "k"), (3, "kc"), (2, "full")])
def test_explicit_stages_keep_the_properties_of_the_last_update(J, tmp_path, scheme, dim, props):
    """NonlinearConductionOperator::Mult (nl_conduction_operator.cpp:281-297) never refreshes the property
    GridFunctions: RK4 stages 2-4 take the gradient of the stage vector with k, rho, C frozen at the vector of the last
    UpdateMatProps (jots_driver.cpp:656-671).  Polynomial properties, so a kernel that re-evaluated them on the stage
    vector differs from the oracle by 1e-4, far above the tolerance."""
    dev, orc = build(J, tmp_path, "Nonlinear_Unsteady", dim, 2, props, scheme=scheme, dt="1e-2", steps=6, n=(4, 3, 2), shear=0.0,
                     bcs=["HeatFlux, 7.5e5", "Isothermal, 350"] + [HF0] * (1 if dim == 2 else 3) + ["HeatFlux, -2e5"])
    assert dev.run() == 6
    u_ref = orc.run()   # refreshing the properties on every stage would move this field by 7e-5 .. 3e-4 (relative l2)
    assert rel(dev.get_u(), u_ref) < FIELD_TOL


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_regressions_against_committed_golden_fields(J, tmp_path, name):
    """Device field vs the committed oracle fixture (tests/golden/regressions.npz, made by
    tests/golden/make_golden.py) for all eight regression configurations: <= 1e-8 relative l2."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "regressions.npz"))
    dev = J.Driver(make_case(name, tmp_path))
    dev.run()
    assert rel(dev.get_u(), gold[name]) < FIELD_TOL
    st = dev.ctx.stats()
    assert st["steps"] == gold[name + "__stats"][0]
    assert abs(st["newton_iters"] - gold[name + "__stats"][1]) <= 2


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_regressions(J, tmp_path, name):
    """The reference's own regression cases: device field vs the oracle's (<= 1e-8 relative l2) and
    vs the analytic solution with the reference's error functional and tolerance."""
    from oracle import analytic, jots as OJ
    ini = make_case(name, tmp_path)
    dev = J.Driver(ini)
    dev.run()
    u = dev.get_u()
    cfg = OJ.Config(ini)
    orc = OJ.JOTSDriver(cfg)
    t_end = cfg.max_timesteps * cfg.dt if cfg.using_time_integration else 0.0
    err = analytic.reference_error(orc.space, u, getattr(analytic, CASES[name]["exact"]), t_end)
    assert np.isfinite(err) and err <= CASES[name]["eps"], (name, err)
    if name in ("Reinert_B1", "NL_Reinert_B2", "Danish_Model1", "Danish_Model1_Neumann"):
        assert rel(u, orc.run()) < FIELD_TOL


def test_host_and_device_pointers_agree(J, tmp_path):
    import torch
    dev, orc = build(J, tmp_path, "Linearized_Unsteady", 3, 2, "k")
    n = dev.n
    x = rnd(n, 7)
    state = smooth_state(orc)
    y_h = dev.op.form_apply(J.FORM_SYSTEM, x, np.empty(n), state=state)
    xd = torch.tensor(x, device="cuda")
    sd = torch.tensor(state, device="cuda")
    yd = torch.empty(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    dev.op.form_apply(J.FORM_SYSTEM, xd, yd, state=sd)
    assert rel(yd.cpu().numpy(), y_h) < 1e-14


def test_errors_are_reported_not_fatal(J, tmp_path):
    bad = tmp_path / "bad.ini"
    bad.write_text("[FiniteElementSetup]\nMesh_File=/nonexistent.mesh\n[MaterialProperties]\n[BoundaryConditions]\n")
    with pytest.raises(J.JotsError):
        J.Driver(str(bad))


# ---- the shipped 2-D example configurations, restated on regenerated meshes --------------------------
_EX = {
    # examples/steady_heat/Heated_Plate/Heated_Plate.ini: [NewtonSolverSettings] is commented out ->
    # newton_max_iter = 0 -> zero Newton iterations; "on" = section enabled
    "Heated_Plate_steady_as_shipped": dict(sim="Steady", order=2, mesh=("plate", 12, 6), t0=310, newton=None, time=False,
        props=[("Thermal_Conductivity_k", "Uniform, 100")], bcs=["Isothermal, 310", HF0, "Isothermal, 300", HF0]),
    "Heated_Plate_steady_newton_on": dict(sim="Steady", order=2, mesh=("plate", 12, 6), t0=310, newton=20, time=False,
        props=[("Thermal_Conductivity_k", "Uniform, 100")], bcs=["Isothermal, 310", HF0, "Isothermal, 300", HF0]),
    # examples/linearized_unsteady_heat/Sinusoidal_Isothermal_Heated_Bar (Q1, bar.mesh, 3 attributes)
    "Sinusoidal_Isothermal_Heated_Bar": dict(sim="Linearized_Unsteady", order=1, mesh=("bar", 40, 4), t0=300, steps=60, dt="0.001",
        props=[("Density_rho", "Uniform, 0.1"), ("Specific_Heat_C", "Uniform, 1000"), ("Thermal_Conductivity_k", "Uniform, 100")],
        bcs=[HF0, "Isothermal, 300", "Sinusoidal_Isothermal, 5, 6.283185307, 0, 300"]),
    "Sinusoidal_HeatFlux_Heated_Bar": dict(sim="Linearized_Unsteady", order=1, mesh=("bar", 40, 4), t0=300, steps=60, dt="0.001",
        props=[("Density_rho", "Uniform, 0.1"), ("Specific_Heat_C", "Uniform, 1000"), ("Thermal_Conductivity_k", "Uniform, 100")],
        bcs=[HF0, "Isothermal, 300", "Sinusoidal_HeatFlux, 1000, 6.283185307, 0, 0"]),
    # examples/linearized_unsteady_heat/Quiescent_IC_HeatFlux: explicit Euler, Max_Timesteps=5e6 is not
    # an int -> boost falls back to the default 100 steps
    "Quiescent_IC_HeatFlux": dict(sim="Linearized_Unsteady", order=2, mesh=("plate", 8, 4), t0=300, steps="5e6", dt="1e-7",
        scheme="Euler_Explicit", props=[("Density_rho", "Uniform, 1"), ("Specific_Heat_C", "Uniform, 1"), ("Thermal_Conductivity_k", "Uniform, 0.5")],
        bcs=[HF0, HF0, HF0, HF0]),
}


@pytest.mark.parametrize("name", sorted(_EX))
def test_shipped_2d_example_configs(J, tmp_path, name):
    from oracle import jots as OJ
    c = _EX[name]
    mesh = str(tmp_path / "m.mesh")
    kind, nx, ny = c["mesh"]
    if kind == "plate":
        write_quad_mesh(mesh, nx, ny, 1.0, 0.25, y0=-0.25)
    else:
        write_quad_mesh(mesh, nx, ny, 1.0, 0.1, attrs=(1, 3, 1, 2))
    txt = ini_text(c["sim"], mesh, c["t0"], c["props"], c["bcs"], time=c.get("time", True), newton=c.get("newton"),
                   steps=1, dt=c.get("dt", "1e-2"), order=c["order"], scheme=c.get("scheme", "Euler_Implicit"))
    txt = txt.replace("Max_Timesteps=1\n", "Max_Timesteps=%s\n" % c.get("steps", 1))
    ini = str(tmp_path / "c.ini")
    with open(ini, "w") as f:
        f.write(txt)
    dev = J.Driver(ini)
    orc = OJ.JOTSDriver(OJ.Config(ini))
    n_dev = dev.run()
    u_ref = orc.run()
    assert n_dev == orc.it.stats["steps"]
    if name == "Quiescent_IC_HeatFlux":
        assert n_dev == 100
    if name == "Heated_Plate_steady_as_shipped":
        assert dev.ctx.stats()["newton_iters"] == 0
    assert rel(dev.get_u(), u_ref) < FIELD_TOL


def test_blow_up_is_reported_like_the_reference(J, tmp_path):
    """An unstable explicit run trips the u.Max() > 1e10 check of PostprocessIteration
    (jots_driver.cpp:739-743): an error status on our side, MFEM_ABORT in the reference."""
    dev, orc = build(J, tmp_path, "Linearized_Unsteady", 2, 2, "const", scheme="Euler_Explicit", dt="1e-1", steps=400, n=(6, 4, 1))
    with pytest.raises(J.JotsError, match="blown up"):
        dev.run()
    with pytest.raises(RuntimeError, match="blown up"):
        orc.run()


@pytest.mark.parametrize("dim,p", [(3, 1), (3, 2), (3, 3), (2, 2), (2, 3)])
@pytest.mark.parametrize("props", ["k", "full"])
def test_general_multilinear_elements(J, tmp_path, dim, p, props):
    """Non-affine (jittered) hexes / quads: per-quadrature-point Jacobians of the multilinear map."""
    dev, orc = build(J, tmp_path, "Nonlinear_Unsteady", dim, p, props, grade=0.3, shear=0.2, jitter=0.2, n=(4, 3, 3), steps=2)
    n = dev.n
    un = smooth_state(orc)
    k = 50.0 * rnd(n, 21)
    x = rnd(n, 22)
    it = orc.it
    it.u_n = un
    it.new_state(k)
    r = dev.op.form_apply(J.FORM_RESIDUAL, k, np.empty(n), state=un)
    assert rel(r, it.mult(k)) < APPLY_TOL
    Jm = it.gradient(k)
    z = un + it.dt * k
    y = dev.op.form_apply(J.FORM_SYSTEM, x, np.empty(n), state=z)
    assert relb(it.ess, y, Jm @ x) < APPLY_TOL
    d = dev.op.form_diagonal(J.FORM_SYSTEM, np.empty(n), state=z)
    assert relb(it.ess, d, Jm.diagonal()) < APPLY_TOL
    assert rel(dev.op.neumann_vector(), it.b_vec) < APPLY_TOL
    assert dev.run() == 2
    assert rel(dev.get_u(), orc.run()) < FIELD_TOL


@pytest.mark.parametrize("props", ["full", "k", "const"])
def test_material_property_output_fields(J, tmp_path, props):
    """Property fields the reference's OutputManager projects for output (output_manager.cpp:63-73, jots_driver.cpp:343)
    and their T-derivatives (material_property.cpp:36-81): device Horner evaluation vs the oracle's repeated products."""
    dev, orc = build(J, tmp_path, "Nonlinear_Unsteady", 3, 2, props)
    state = smooth_state(orc)
    dev.ctx.set_material_state(state)
    for which, short in enumerate(("rho", "C", "k")):
        mp = orc.props[short]
        mp.update(state)
        for order, ref in enumerate((mp.coeff, mp.dcoeff, mp.d2coeff)):
            f = dev.ctx.material_field(which, order)
            ref = np.broadcast_to(np.asarray(ref, dtype=float), f.shape)
            assert np.abs(f - ref).max() <= 1e-13 * max(np.abs(ref).max(), 1e-300), (short, order)


@pytest.mark.parametrize("sim,props,n,max_it", [("Nonlinear_Unsteady", "k", (6, 3, 2), 100), ("Linearized_Unsteady", "const", (9, 5, 3), 100),
                                                ("Nonlinear_Unsteady", "k", (8, 8, 4), 7), ("Steady", "k", (6, 3, 2), 100)])
def test_cg_with_device_resident_scalars_matches_the_host_sequenced_loop(J, tmp_path, monkeypatch, sim, props, n, max_it):
    """CG + Jacobi keeps alpha, beta and the stopping test on the device and enqueues iterations ahead of the stop flag
    (solvers.cpp: cg_device_loop); kernels past the stopping iteration must be no-ops: same Krylov / Newton counts and the same
    field as the loop that reads the scalars back every iteration (mfem::CGSolver::Mult control flow), whether the solve
    converges early, stops at an iteration that is not a multiple of the chunk, or runs into Max_Iterations."""
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("JOTS_CG_HOST_SCALARS", mode)
        dev, orc = build(J, tmp_path, sim, 3, 2, props, solver="CG", prec="Jacobi", n=n, steps=2)
        if max_it != 100:
            # fewer Krylov iterations than a chunk: Newton may not converge, the comparison is between the two loops only
            txt = open(str(tmp_path / "c.ini")).read().replace("Max_Iterations=100\n", "Max_Iterations=%d\n" % max_it, 1)   # [LinearSolverSettings]
            open(str(tmp_path / "c.ini"), "w").write(txt)
            dev = J.Driver(str(tmp_path / "c.ini"))
        try:
            dev.run()
        except J.JotsError:
            pass      # Newton non-convergence with a 7-iteration Krylov budget is fatal, as in the reference
        st = dev.ctx.stats()
        out[mode] = (st["krylov_iters"], st["newton_iters"], st["last_linear_iters"], st["last_linear_converged"], dev.get_u())
    assert out["0"][:4] == out["1"][:4], (out["0"][:4], out["1"][:4])
    assert rel(out["0"][4], out["1"][4]) < 1e-12
    if max_it == 100:
        assert rel(out["0"][4], orc.run()) < FIELD_TOL


def test_print_levels_and_step_lines(J, tmp_path, capfd):
    """Print_Level lists (Factory::CreatePrintLevel, jots_common.inl:57-91; config_file.cpp:140-141,161-162) and Print_Freq
    (jots_driver.cpp:731-736): 'All' prints Newton / Krylov iteration lines and a step line per step, 'None' prints nothing,
    an unknown label is an error (Print_Level_Map.at throws in the reference).  Results do not depend on the print level
    (CG falls back to the host-sequenced loop when it has to print every iteration)."""
    from cases import ini_text
    fields = {}
    for name, lsp, nwp, freq in (("all", "All", "Errors, Warnings, Iterations", 1), ("none", "None", "None", 1000000)):
        mesh = J.Mesh.brick(6, 3, 2, 0.01, 0.01, 0.01)
        ini = str(tmp_path / (name + ".ini"))
        with open(ini, "w") as f:
            f.write(ini_text("Nonlinear_Unsteady", "unused", 300, PROPS["k"], ["HeatFlux, 7.5e5", HF0, HF0, "Isothermal, 300", HF0, HF0],
                             newton=50, solver="CG", prec="Jacobi", steps=2, order=2, ls_print=lsp, newton_print=nwp, print_freq=freq))
        capfd.readouterr()
        dev = J.Driver(ini, 0, mesh=mesh)
        dev.run()
        out = capfd.readouterr().out
        fields[name] = dev.get_u()
        if name == "all":
            assert out.count("Step #") == 2 and "Newton iteration  0 : ||r|| = " in out and "(B r, r) = " in out and "PCG: Number of iterations:" in out
        else:
            assert out.strip() == ""
    assert rel(fields["all"], fields["none"]) < 1e-12
    bad = str(tmp_path / "bad.ini")
    with open(bad, "w") as f:
        f.write(ini_text("Nonlinear_Unsteady", "unused", 300, PROPS["k"], [HF0] * 6, newton=5, steps=1, ls_print="Errors, Chatty"))
    with pytest.raises(J.JotsError, match="print level"):
        J.Driver(bad, 0, mesh=J.Mesh.brick(2, 2, 2, 1.0, 1.0, 1.0))


@pytest.mark.parametrize("solver", ["GMRES", "FGMRES"])
def test_one_block_gram_schmidt_sweep_matches_the_chained_launches(J, tmp_path, solver):
    """Small single-GPU problems run the modified Gram-Schmidt sweep of an (F)GMRES iteration in one block
    (vec.cu: k_mgs_sweep_small); large and partitioned ones chain one launch per basis vector.  Same operation order:
    same iteration counts and fields.  JOTS_MGS_SMALL is read once per process, so the chained path runs in a child."""
    import subprocess, sys, os, json
    dev, orc = build(J, tmp_path, "Nonlinear_Unsteady", 3, 2, "kc", solver=solver, prec="Chebyshev", n=(6, 3, 2), steps=3)
    dev.run()
    st = dev.ctx.stats()
    u = dev.get_u()
    assert rel(u, orc.run()) < FIELD_TOL
    code = ("import sys, json, numpy as np; sys.path.insert(0, %r); import jots_b200 as J; d = J.Driver(%r); d.run(); s = d.ctx.stats(); "
            "np.save(%r, d.get_u()); print(json.dumps([s['krylov_iters'], s['newton_iters'], s['kernel_launches']]))"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(tmp_path / "c.ini"), str(tmp_path / "u.npy")))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, JOTS_MGS_SMALL="0"), capture_output=True, text=True, check=True).stdout
    k2, n2, launches2 = json.loads(out.strip().splitlines()[-1])
    assert (k2, n2) == (st["krylov_iters"], st["newton_iters"])
    assert launches2 > st["kernel_launches"]                      # the chained path launches more kernels
    assert rel(np.load(str(tmp_path / "u.npy")), u) < 1e-12


@pytest.mark.parametrize("p", [1, 2, 3])
def test_device_jacobian_against_exact_integrals(J, tmp_path, p):
    """The device Jacobian-vector product and diagonal of BASELINE configs[3]'s physics (k linear in T, uniform rho C) against
    the matrix assembled from EXACT element integrals (tests/test_oracle_exact.py: mpmath, no quadrature, none of the oracle's
    tables) on a uniform brick -- for hexes of order <= 3 the reference's rule integrates these forms exactly, so this is the
    reference's discrete operator.  Q2: the two-phase patch kernel; Q1, Q3: the gather-table kernel."""
    from test_oracle_exact import element_operators, exact_1d
    n = (5, 4, 3)
    L = (0.01, 0.02, 0.015)
    dev, orc = build(J, tmp_path, "Nonlinear_Unsteady", 3, p, "k", n=n, steps=1)
    s, it = orc.space, orc.it
    h = tuple(L[a] / n[a] for a in range(3))
    u_n = smooth_state(orc)
    kvec = 50.0 * rnd(dev.n, 41)
    z = u_n + it.dt * kvec
    rc = 8000.0 * 500.0
    Jx = np.zeros((s.ndof, s.ndof))
    X = s.node_coordinates()
    for e in range(s.gather.shape[0]):
        g = s.gather[e]
        # local order of the element's nodes: x fastest, then y, then z (integer keys: equal coordinates differ in the last bit)
        key = np.rint((X[g] - X[g].min(0)) / np.array(h) * 1e6).astype(np.int64)
        gl = g[np.lexsort((key[:, 0], key[:, 1], key[:, 2]))]
        Me, Ke, We = element_operators(p, h, 0.09 * z[gl] - 17.0, z[gl], exact_1d(p))
        Jx[np.ix_(gl, gl)] += Me + it.dt * (Ke + 0.09 * We) / rc
    ess = np.asarray(it.ess)
    Jx[ess, :] = 0.0
    Jx[:, ess] = 0.0
    Jx[ess, ess] = 1.0
    free = np.setdiff1d(np.arange(s.ndof), ess)
    x = rnd(dev.n, 42)
    y = dev.op.form_apply(J.FORM_SYSTEM, x, np.empty(dev.n), state=z)
    ref = Jx @ x
    # the unit-diagonal rows (O(1)) dwarf the free rows (1e-9): compare them separately
    assert rel(y[free], ref[free]) < APPLY_TOL, rel(y[free], ref[free])
    assert np.array_equal(y[ess], x[ess])
    d = dev.op.form_diagonal(J.FORM_SYSTEM, np.empty(dev.n), state=z)
    assert rel(d[free], np.diag(Jx)[free]) < APPLY_TOL and np.array_equal(d[ess], np.ones(ess.size))
