"""Oracle (test infrastructure): Krylov / Newton / smoothers with MFEM 4.7 semantics.

The reference builds these through ``Factory::GetSolver`` / ``GetPrec``
(/root/reference/src/jots_common.inl:7-38): ``mfem::CGSolver``, ``GMRESSolver`` (m=50),
``FGMRESSolver`` (m=50), ``HypreSmoother`` Jacobi / Chebyshev, ``NewtonSolver``.  MFEM and
hypre are not vendored in /root/reference; the control flow and stopping rules below
restate MFEM 4.7 ``linalg/solvers.cpp`` and hypre's ``par_cheby.c`` as summarised in
SURVEY.md App. A.4-A.5 (parity unpinned beyond the analytic tests).

All solvers run with ``iterative_mode = false`` (x0 = 0), as every call site in the
reference sets it (linear_conduction_operator.cpp:38,49; nl_conduction_operator.cpp:246,260).
"""
from __future__ import annotations

import numpy as np


class SolveInfo:
    def __init__(self):
        self.converged = False
        self.iterations = 0
        self.final_norm = 0.0
        self.applies = 0  # operator applications (for throughput accounting)


# ----------------------------------------------------------------------------------------------
# preconditioners (HypreSmoother)
# ----------------------------------------------------------------------------------------------

def hash_unit(n, offset=0):
    """Deterministic pseudo-random vector in [-1,1): splitmix64 of the (global) index, so the
    same start vector is obtained on any partition and on the device."""
    z = (np.arange(n, dtype=np.uint64) + np.uint64(offset) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0


class JacobiSmoother:
    """HypreSmoother::Jacobi, one sweep from zero: z = D^-1 r."""

    def __init__(self):
        self.dinv = None

    def set_operator(self, A, diag):
        self.dinv = 1.0 / diag

    def mult(self, r):
        return self.dinv * r


class ChebyshevSmoother:
    """HypreSmoother::Chebyshev (type 16), poly_order 2, poly_fraction 0.3, 10 CG iterations for the
    eigenvalue estimate of D^-1/2 A D^-1/2 (App. A.5).  hypre_ParCSRRelax_Cheby_Setup/Solve use
    ``cheby_order = order - 1`` ("we are using the order of p(A)"), so MFEM's poly_order = 2 is the
    degree-1 polynomial c0 + c1 A^: ONE operator application per smoother call."""

    def __init__(self, poly_order=2, fraction=0.3, eig_iters=10):
        self.order = poly_order - 1          # degree of p
        self.fraction = fraction
        self.eig_iters = eig_iters

    def set_operator(self, A, diag):
        self.A = A
        self.ds = 1.0 / np.sqrt(diag)
        lmax, lmin = self._eig_estimate()
        self.lmax, self.lmin = lmax, lmin
        ub = 1.1 * lmax
        lb = lmin + self.fraction * (ub - lmin)
        theta = 0.5 * (ub + lb)
        delta = 0.5 * (ub - lb)
        if self.order == 0:
            self.c = np.array([1.0 / theta])
        elif self.order == 1:
            den = theta * theta + delta * theta
            self.c = np.array([(delta + 2 * theta) / den, -1.0 / den])
        elif self.order == 2:
            den = 2 * delta * theta ** 2 - delta ** 2 * theta - delta ** 3 + 2 * theta ** 3
            self.c = np.array([(4 * delta * theta - delta ** 2 + 6 * theta ** 2) / den,
                               -(2 * delta + 6 * theta) / den,
                               2.0 / den])
        else:
            raise ValueError("poly_order > 3 is not used by JOTS (HypreSmoother default is 2)")

    def _ahat(self, v):
        return self.ds * (self.A @ (self.ds * v))

    def _eig_estimate(self):
        """CG-Lanczos on the scaled operator from a deterministic pseudo-random vector."""
        n = self.ds.size
        r = hash_unit(n)
        p = r.copy()
        rr = r @ r
        alphas, betas = [], []
        for _ in range(self.eig_iters):
            Ap = self._ahat(p)
            pAp = p @ Ap
            if pAp <= 0 or rr == 0:
                break
            alpha = rr / pAp
            r = r - alpha * Ap
            rr_new = r @ r
            beta = rr_new / rr
            alphas.append(alpha)
            betas.append(beta)
            p = r + beta * p
            rr = rr_new
        return lanczos_extreme_eigs(alphas, betas)

    def mult(self, r):
        rs = self.ds * r
        u = self.c[self.order] * rs
        for i in range(self.order - 1, -1, -1):
            u = self._ahat(u) + self.c[i] * rs
        return self.ds * u


def lanczos_extreme_eigs(alphas, betas):
    """Extreme eigenvalues of the Lanczos tridiagonal built from CG coefficients."""
    m = len(alphas)
    if m == 0:
        return 1.0, 1.0
    T = np.zeros((m, m))
    for j in range(m):
        T[j, j] = 1.0 / alphas[j] + (betas[j - 1] / alphas[j - 1] if j > 0 else 0.0)
        if j + 1 < m:
            T[j, j + 1] = T[j + 1, j] = np.sqrt(betas[j]) / alphas[j]
    ev = np.linalg.eigvalsh(T)
    return float(ev[-1]), float(ev[0])


# ----------------------------------------------------------------------------------------------
# Krylov
# ----------------------------------------------------------------------------------------------

class _Krylov:
    def __init__(self, rel_tol=1e-10, abs_tol=1e-16, max_iter=100):
        self.rel_tol, self.abs_tol, self.max_iter = rel_tol, abs_tol, max_iter
        self.prec = None
        self.A = None
        self.info = SolveInfo()

    def set_operator(self, A, diag=None):
        self.A = A
        if self.prec is not None:
            self.prec.set_operator(A, A.diagonal() if diag is None else diag)

    def _apply(self, x):
        self.info.applies += 1
        return self.A @ x


class CGSolver(_Krylov):
    """mfem::CGSolver::Mult, preconditioned, monitors (B r, r)."""

    def mult(self, b):
        info = self.info = SolveInfo()
        x = np.zeros_like(b)
        r = b.copy()
        z = self.prec.mult(r) if self.prec else r.copy()
        d = z.copy()
        nom = d @ r
        if nom < 0:
            return x
        r0 = max(nom * self.rel_tol ** 2, self.abs_tol ** 2)
        if nom <= r0:
            info.converged, info.final_norm = True, np.sqrt(nom)
            return x
        z = self._apply(d)
        den = z @ d
        if den == 0.0:
            info.final_norm = np.sqrt(nom)
            return x
        info.iterations = self.max_iter
        i = 1
        while True:
            alpha = nom / den
            x += alpha * d
            r -= alpha * z
            if self.prec:
                z = self.prec.mult(r)
                betanom = r @ z
            else:
                betanom = r @ r
            if betanom < 0:
                info.iterations = i
                break
            if betanom <= r0:
                info.converged, info.iterations = True, i
                break
            i += 1
            if i > self.max_iter:
                break
            beta = betanom / nom
            d = (z if self.prec else r) + beta * d
            z = self._apply(d)
            den = d @ z
            if den == 0.0:
                info.iterations = i
                break
            nom = betanom
        info.final_norm = np.sqrt(abs(betanom))
        return x


def _gen_rot(dx, dy):
    if dy == 0.0:
        return 1.0, 0.0
    if abs(dy) > abs(dx):
        t = dx / dy
        sn = 1.0 / np.sqrt(1.0 + t * t)
        return t * sn, sn
    t = dy / dx
    cs = 1.0 / np.sqrt(1.0 + t * t)
    return cs, t * cs


def _app_rot(dx, dy, cs, sn):
    return cs * dx + sn * dy, -sn * dx + cs * dy


def _update(x, k, H, s, V):
    y = s[:k + 1].copy()
    for i in range(k, -1, -1):
        y[i] /= H[i, i]
        for j in range(i - 1, -1, -1):
            y[j] -= H[j, i] * y[i]
    for j in range(k + 1):
        x += y[j] * V[j]


class FGMRESSolver(_Krylov):
    """mfem::FGMRESSolver::Mult (m=50): right preconditioning, true residual norm."""
    m = 50

    def mult(self, b):
        info = self.info = SolveInfo()
        m = self.m
        x = np.zeros_like(b)
        r = b.copy()
        beta = np.linalg.norm(r)
        final_norm = max(self.rel_tol * beta, self.abs_tol)
        if beta <= final_norm:
            info.converged, info.final_norm = True, beta
            return x
        H = np.zeros((m + 1, m))
        s = np.zeros(m + 1)
        cs = np.zeros(m + 1)
        sn = np.zeros(m + 1)
        V = [None] * (m + 1)
        Z = [None] * (m + 1)
        j = 1
        while j <= self.max_iter:
            V[0] = r / beta
            s[:] = 0.0
            s[0] = beta
            i = 0
            while i < m and j <= self.max_iter:
                Z[i] = self.prec.mult(V[i]) if self.prec else V[i].copy()
                r = self._apply(Z[i])
                for k in range(i + 1):
                    H[k, i] = r @ V[k]
                    r = r - H[k, i] * V[k]
                H[i + 1, i] = np.linalg.norm(r)
                V[i + 1] = r / H[i + 1, i]
                for k in range(i):
                    H[k, i], H[k + 1, i] = _app_rot(H[k, i], H[k + 1, i], cs[k], sn[k])
                cs[i], sn[i] = _gen_rot(H[i, i], H[i + 1, i])
                H[i, i], H[i + 1, i] = _app_rot(H[i, i], H[i + 1, i], cs[i], sn[i])
                s[i], s[i + 1] = _app_rot(s[i], s[i + 1], cs[i], sn[i])
                resid = abs(s[i + 1])
                if resid <= final_norm:
                    _update(x, i, H, s, Z)
                    info.converged, info.final_norm, info.iterations = True, resid, j
                    return x
                i += 1
                j += 1
            _update(x, i - 1, H, s, Z)
            r = b - self._apply(x)
            beta = np.linalg.norm(r)
            if beta <= final_norm:
                info.converged, info.final_norm, info.iterations = True, beta, j
                return x
        info.final_norm, info.iterations = beta, self.max_iter
        return x


class GMRESSolver(_Krylov):
    """mfem::GMRESSolver::Mult (m=50): left preconditioning, monitors ||B r||."""
    m = 50

    def mult(self, b):
        info = self.info = SolveInfo()
        m = self.m
        x = np.zeros_like(b)
        P = (lambda v: self.prec.mult(v)) if self.prec else (lambda v: v.copy())
        r = P(b)
        beta = np.linalg.norm(r)
        final_norm = max(self.rel_tol * beta, self.abs_tol)
        if beta <= final_norm:
            info.converged, info.final_norm = True, beta
            return x
        H = np.zeros((m + 1, m))
        s = np.zeros(m + 1)
        cs = np.zeros(m + 1)
        sn = np.zeros(m + 1)
        V = [None] * (m + 1)
        j = 1
        while j <= self.max_iter:
            V[0] = r / beta
            s[:] = 0.0
            s[0] = beta
            i = 0
            while i < m and j <= self.max_iter:
                w = P(self._apply(V[i]))
                for k in range(i + 1):
                    H[k, i] = w @ V[k]
                    w = w - H[k, i] * V[k]
                H[i + 1, i] = np.linalg.norm(w)
                V[i + 1] = w / H[i + 1, i]
                for k in range(i):
                    H[k, i], H[k + 1, i] = _app_rot(H[k, i], H[k + 1, i], cs[k], sn[k])
                cs[i], sn[i] = _gen_rot(H[i, i], H[i + 1, i])
                H[i, i], H[i + 1, i] = _app_rot(H[i, i], H[i + 1, i], cs[i], sn[i])
                s[i], s[i + 1] = _app_rot(s[i], s[i + 1], cs[i], sn[i])
                resid = abs(s[i + 1])
                if resid <= final_norm:
                    _update(x, i, H, s, V)
                    info.converged, info.final_norm, info.iterations = True, resid, j
                    return x
                i += 1
                j += 1
            _update(x, i - 1, H, s, V)
            r = P(b - self._apply(x))
            beta = np.linalg.norm(r)
            if beta <= final_norm:
                info.converged, info.final_norm, info.iterations = True, beta, j
                return x
        info.final_norm, info.iterations = beta, self.max_iter
        return x


def make_solver(label, rel_tol, abs_tol, max_iter):
    cls = {"CG": CGSolver, "GMRES": GMRESSolver, "FGMRES": FGMRESSolver}[label]
    return cls(rel_tol, abs_tol, max_iter)


def make_prec(label):
    return {"Jacobi": JacobiSmoother, "Chebyshev": ChebyshevSmoother}[label]()


# ----------------------------------------------------------------------------------------------
# Newton
# ----------------------------------------------------------------------------------------------

class NewtonSolver:
    """mfem::NewtonSolver::Mult with the JOTSNewtonSolver::ProcessNewState hook
    (/root/reference/src/jots_common.inl:124-147).  ``oper`` provides ``mult(x)``,
    ``gradient(x)`` -> (csr matrix) and ``new_state(x)``."""

    def __init__(self, rel_tol=1e-10, abs_tol=1e-16, max_iter=100):
        self.rel_tol, self.abs_tol, self.max_iter = rel_tol, abs_tol, max_iter
        self.converged = False
        self.iterations = 0
        self.final_norm = 0.0
        self.linear_iterations = 0
        self.linear_applies = 0
        self.residual_evals = 0

    def mult(self, oper, lin_solver, b, x):
        oper.new_state(x)
        r = oper.mult(x)
        self.residual_evals = 1
        if b is not None:
            r = r - b
        norm = np.linalg.norm(r)
        goal = max(self.rel_tol * norm, self.abs_tol)
        it = 0
        self.linear_iterations = 0
        self.linear_applies = 0
        while True:
            if norm <= goal:
                self.converged = True
                break
            if it >= self.max_iter:
                self.converged = False
                break
            J = oper.gradient(x)
            lin_solver.set_operator(J)
            c = lin_solver.mult(r)
            self.linear_iterations += lin_solver.info.iterations
            self.linear_applies += lin_solver.info.applies
            x = x - c
            oper.new_state(x)
            r = oper.mult(x)
            self.residual_evals += 1
            if b is not None:
                r = r - b
            norm = np.linalg.norm(r)
            it += 1
        self.iterations = it
        self.final_norm = norm
        return x
