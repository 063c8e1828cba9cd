"""The oracle's Krylov solvers (oracle/solvers.py: restatements of mfem::CGSolver / GMRESSolver / FGMRESSolver, SURVEY.md App.
A.4) against their DEFINITIONS, computed densely and independently of the recurrences: after k iterations from x0 = 0

  * preconditioned CG returns the minimiser of the A-norm of the error over  B K_k(A B, b),
  * GMRES with left preconditioning the minimiser of ||B (b - A x)|| over  K_k(B A, B b),
  * FGMRES (fixed B: right preconditioning) the minimiser of ||b - A x|| over  B K_k(A B, b),

and the (B r, r) / residual norms they report are the norms of those minimisers.  CG is also compared with SciPy's
implementation (same iterates).  The systems are the oracle's own eliminated operators (linearized A = M + dt K on a small
brick; the non-symmetric Newton Jacobian with k(T) for the GMRES variants), the preconditioner is Jacobi.  This pins the
arithmetic of the solvers that the device solvers are compared with; MFEM's stopping rules themselves stay as recollected."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cases import ini_text  # noqa: E402
from oracle import jots as OJ, mesh as OM, solvers as OS  # noqa: E402

PROPS = [("Density_rho", "Uniform, 8000"), ("Specific_Heat_C", "Uniform, 500"), ("Thermal_Conductivity_k", "Polynomial, 0.09, -17")]
BCS = ["HeatFlux, 7.5e5", "HeatFlux, 0", "HeatFlux, 0", "Isothermal, 300", "HeatFlux, 0", "HeatFlux, 0"]


def system(tmp_path, sim):
    ini = str(tmp_path / "c.ini")
    with open(ini, "w") as f:
        f.write(ini_text(sim, "unused.mesh", 300, PROPS, BCS, newton=None if sim == "Linearized_Unsteady" else 50, solver="CG", prec="Jacobi",
                         steps=1, order=2))
    orc = OJ.JOTSDriver(OJ.Config(ini), mesh=OM.make_brick(4, 3, 2, 0.01, 0.01, 0.01))
    rng = np.random.default_rng(3)
    n = orc.space.ndof
    if sim == "Linearized_Unsteady":
        orc.u = 300.0 + 400.0 * rng.random(n)
        orc.update_mat_props(True)
        A = orc.it.A_mat
    else:
        it = orc.it
        it.u_n = 300.0 + 400.0 * rng.random(n)
        k = 1000.0 * (rng.random(n) - 0.5)
        it.new_state(k)
        A = it.gradient(k)
    b = rng.standard_normal(n) * A.diagonal()          # entries of the scale of the rows
    b[np.asarray(orc.it.ess)] = 0.0
    return A.tocsr(), b


def krylov_basis(op, v0, k):
    """orthonormal basis of span{v0, op v0, ..., op^(k-1) v0} (full re-orthogonalisation, float64)"""
    Q = np.zeros((v0.size, k))
    w = v0.copy()
    for j in range(k):
        for _ in range(2):
            w = w - Q[:, :j] @ (Q[:, :j].T @ w)
        Q[:, j] = w / np.linalg.norm(w)
        w = op(Q[:, j])
    return Q


@pytest.mark.parametrize("k", [1, 5, 12])
def test_cg_iterates_minimise_the_energy_norm_and_match_scipy(tmp_path, k):
    A, b = system(tmp_path, "Linearized_Unsteady")
    dinv = 1.0 / A.diagonal()
    s = OS.CGSolver(rel_tol=0.0, abs_tol=0.0, max_iter=k)
    s.prec = OS.JacobiSmoother()
    s.set_operator(A)
    x = s.mult(b)
    assert s.info.iterations == k and not s.info.converged
    # definition: x_k = argmin ||x - x*||_A over B K_k(A B, b)  <=>  Galerkin: W^T A W y = W^T b
    W = dinv[:, None] * krylov_basis(lambda v: A @ (dinv * v), b, k)
    y = np.linalg.solve(W.T @ (A @ W), W.T @ b)
    ref = W @ y
    assert np.linalg.norm(x - ref) <= 1e-9 * np.linalg.norm(ref)
    r = b - A @ x
    assert abs(s.info.final_norm - np.sqrt(r @ (dinv * r))) <= 1e-9 * s.info.final_norm          # sqrt((B r, r))
    # SciPy's preconditioned CG produces the same iterates
    xs, _ = spla.cg(A, b, x0=np.zeros_like(b), M=spla.LinearOperator(A.shape, matvec=lambda v: dinv * v), maxiter=k, rtol=0.0, atol=0.0)
    assert np.linalg.norm(x - xs) <= 1e-9 * np.linalg.norm(x)


@pytest.mark.parametrize("k", [1, 6, 14])
@pytest.mark.parametrize("flexible", [False, True])
def test_gmres_iterates_minimise_their_residual_norms(tmp_path, k, flexible):
    A, b = system(tmp_path, "Nonlinear_Unsteady")
    assert abs(A - A.T).max() > 0          # the Newton Jacobian with k(T) is not symmetric
    dinv = 1.0 / A.diagonal()
    s = (OS.FGMRESSolver if flexible else OS.GMRESSolver)(rel_tol=0.0, abs_tol=0.0, max_iter=k)
    s.prec = OS.JacobiSmoother()
    s.set_operator(A)
    x = s.mult(b)
    assert s.info.iterations == k and not s.info.converged
    if flexible:
        # min ||b - A x|| over x in B K_k(A B, b)
        W = dinv[:, None] * krylov_basis(lambda v: A @ (dinv * v), b, k)
        y = np.linalg.lstsq(A @ W, b, rcond=None)[0]
        norm = np.linalg.norm(b - A @ (W @ y))
    else:
        # min ||B (b - A x)|| over x in K_k(B A, B b)
        W = krylov_basis(lambda v: dinv * (A @ v), dinv * b, k)
        y = np.linalg.lstsq(dinv[:, None] * (A @ W), dinv * b, rcond=None)[0]
        norm = np.linalg.norm(dinv * (b - A @ (W @ y)))
    ref = W @ y
    assert np.linalg.norm(x - ref) <= 1e-8 * np.linalg.norm(ref)
    assert abs(s.info.final_norm - norm) <= 1e-8 * norm


def test_chebyshev_smoother_is_hypres_degree_one_polynomial(tmp_path):
    """HypreSmoother::Chebyshev as MFEM sets it up (type 16, poly_order 2 -> hypre's cheby_order 1, fraction 0.3): one
    application is z = D^-1/2 p(A^) D^-1/2 r with A^ = D^-1/2 A D^-1/2 and hypre's p(t) = (delta - t + 2 theta) / (theta^2 +
    delta theta) (par_cheby.c, case 1), whose residual polynomial 1 - t p(t) has its roots at theta and at the upper bound
    1.1 lambda_max; evaluated here through a dense eigendecomposition instead of the operator recurrence."""
    A, b = system(tmp_path, "Linearized_Unsteady")
    sm = OS.ChebyshevSmoother()
    sm.set_operator(A, A.diagonal())
    ds = 1.0 / np.sqrt(A.diagonal())
    Ah = (ds[:, None] * A.toarray()) * ds[None, :]
    lam, Q = np.linalg.eigh(0.5 * (Ah + Ah.T))
    # the Lanczos estimate after 10 steps brackets nothing exactly, but must sit inside the spectrum and near its top
    assert lam[0] - 1e-12 <= sm.lmin <= sm.lmax <= lam[-1] + 1e-12 and sm.lmax > 0.9 * lam[-1]
    ub = 1.1 * sm.lmax
    lb = sm.lmin + 0.3 * (ub - sm.lmin)
    theta, delta = 0.5 * (ub + lb), 0.5 * (ub - lb)
    p = lambda t: (delta - t + 2.0 * theta) / (theta * theta + delta * theta)
    assert abs(1.0 - theta * p(theta)) < 1e-14 and abs(1.0 - ub * p(ub)) < 1e-14
    z = sm.mult(b)
    ref = ds * (Q @ (p(lam) * (Q.T @ (ds * b))))
    assert np.linalg.norm(z - ref) <= 1e-12 * np.linalg.norm(ref)
    assert sm.order == 1 and sm.c.size == 2
