"""Exact-integral pins of the oracle's element operators (CPU; mpmath, independent of oracle/fem.py's tables).

The reference integrates every form with MFEM's default Gauss rules (p + 2 points per direction on hexes, p + 1 on quads:
SURVEY.md App. A.2).  On affine hexes of order p <= 3 the integrands of the benchmarked physics are polynomials of degree
<= 2p + 3 per direction -- mass (2p), diffusion with the nodal interpolant k_h of k(T) (3p), the extra Jacobian term of
JOTSNonlinearDiffusionIntegrator::AssembleElementGrad with a uniform k' (k linear in T: 3p) -- so the rule is EXACT and the
discrete operator is the exact Galerkin integral, whatever rule of sufficient order a library picks.  These tests integrate the
Lagrange basis on the Gauss-Lobatto points (closed-form nodes) exactly (term-by-term integrals of the product polynomials in 40-digit arithmetic, mpmath) and require
the oracle's matrices to agree to 1e-13: that pins the basis, its nodes, the nodal-interpolation semantics of the coefficients
and the tensor assembly below the reference's analytic tolerances, independently of the oracle's own tables.  Where the rule is
NOT exact (quads use p + 1 points: k_h-weighted diffusion at p >= 2) the exact integral must differ visibly and the oracle must
equal the (p + 1)-point Gauss sum instead, formed here from Legendre roots found by Newton iteration in the same arithmetic -- so the quadrature order is pinned too."""
import functools

import mpmath as mp
import numpy as np
import pytest

from oracle import fem, mesh as M

mp.mp.dps = 40


@functools.lru_cache(maxsize=None)
def gll_nodes(p):
    """Gauss-Lobatto points on [0, 1] in closed form: the ends and the roots of P_p'(2x - 1)  (p = 3: +-1/sqrt 5, p = 4: 0, +-sqrt(3/7))."""
    inner = {1: [], 2: [mp.mpf(0)], 3: [-1 / mp.sqrt(5), 1 / mp.sqrt(5)], 4: [-mp.sqrt(mp.mpf(3) / 7), mp.mpf(0), mp.sqrt(mp.mpf(3) / 7)]}[p]
    return tuple([mp.mpf(0)] + [(r + 1) / 2 for r in inner] + [mp.mpf(1)])


def _pmul(a, b):
    out = [mp.mpf(0)] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            out[i + j] += x * y
    return out


def _pder(a):
    return [k * a[k] for k in range(1, len(a))] or [mp.mpf(0)]


def _pint01(a):
    return sum(c / (k + 1) for k, c in enumerate(a))


def _pval(a, x):
    return sum(c * x ** k for k, c in enumerate(a))


@functools.lru_cache(maxsize=None)
def basis(p):
    """Lagrange polynomials on the Gauss-Lobatto points as coefficient lists (40-digit arithmetic)."""
    xs = gll_nodes(p)
    out = []
    for i in range(p + 1):
        f = [mp.mpf(1)]
        for m in range(p + 1):
            if m != i:
                f = _pmul(f, [-xs[m] / (xs[i] - xs[m]), 1 / (xs[i] - xs[m])])
        out.append(f)
    return tuple(tuple(f) for f in out)


@functools.lru_cache(maxsize=None)
def exact_1d(p):
    """A3[a,i,j] = int phi_a phi_i phi_j,  B3[a,i,j] = int phi_a phi_i' phi_j'  on [0, 1]: term-by-term integrals of the
    product polynomials, exact up to the 40-digit arithmetic."""
    ph = [list(f) for f in basis(p)]
    dph = [_pder(f) for f in ph]
    D = p + 1
    A3 = np.zeros((D, D, D))
    B3 = np.zeros((D, D, D))
    for a in range(D):
        for i in range(D):
            for j in range(i, D):
                A3[a, i, j] = A3[a, j, i] = float(_pint01(_pmul(ph[a], _pmul(ph[i], ph[j]))))
                B3[a, i, j] = B3[a, j, i] = float(_pint01(_pmul(ph[a], _pmul(dph[i], dph[j]))))
    return A3, B3


@functools.lru_cache(maxsize=None)
def gauss_1d(p, nq):
    """The same two tensors from an nq-point Gauss-Legendre sum: roots of P_nq by Newton iteration in 40-digit arithmetic,
    weights 2 / ((1 - r^2) P_nq'(r)^2) (mapped to [0, 1])."""
    roots = []
    for k in range(nq):
        r = mp.cos(mp.pi * (k + mp.mpf(3) / 4) / (nq + mp.mpf(1) / 2))
        for _ in range(60):
            r = r - mp.legendre(nq, r) / mp.diff(lambda t: mp.legendre(nq, t), r)
        roots.append(r)
    roots.sort()
    pts = [(r + 1) / 2 for r in roots]
    wts = [1 / ((1 - r ** 2) * mp.diff(lambda t: mp.legendre(nq, t), r) ** 2) for r in roots]
    ph = [list(f) for f in basis(p)]
    dph = [_pder(f) for f in ph]
    v = np.array([[float(_pval(f, q)) for f in ph] for q in pts])
    dv = np.array([[float(_pval(f, q)) for f in dph] for q in pts])
    w = np.array([float(x) for x in wts])
    assert abs(w.sum() - 1.0) < 1e-15
    return np.einsum("q,qa,qi,qj->aij", w, v, v, v), np.einsum("q,qa,qi,qj->aij", w, v, dv, dv)


def element_operators(p, h, kn, zn, tensors):
    """One affine element [0, hx] x [0, hy] (x [0, hz]) with nodal coefficient kn and nodal state zn (lexicographic, x fastest):
    M_ij = int phi_i phi_j,  K_ij = int k_h grad phi_i . grad phi_j,  W_im = int phi_m grad z_h . grad phi_i."""
    dim = len(h)
    A3, B3 = tensors
    D = p + 1
    one = np.ones(D)
    A2 = np.einsum("a,aij->ij", one, A3)          # int phi_i phi_j (the basis sums to one)
    shape = (D,) * dim
    k = kn.reshape(shape[::-1])                   # [c, b, a] = [z, y, x]
    z = zn.reshape(shape[::-1])
    vol = np.prod(h)
    if dim == 3:
        Mm = vol * np.einsum("il,jm,kn->kjinml", A2, A2, A2).reshape(D ** 3, D ** 3)
        K = np.zeros((D,) * 6)
        W = np.zeros((D,) * 6)
        for d in range(3):
            T = [B3 if a == d else A3 for a in range(3)]              # x, y, z factors
            s = vol / h[d] ** 2
            # indices: node n = (c, b, a); rows (k, j, i); cols (n_, m, l) all stored [z, y, x]
            K += s * np.einsum("cba,ail,bjm,ckn->kjinml", k, T[0], T[1], T[2])
            # W_{row i, col m} = sum_n z_n int phi_m d phi_n d phi_i : factor tensor [m, n, i]
            W += s * np.einsum("cba,lai,mbj,nck->kjinml", z, T[0], T[1], T[2])
        return Mm, K.reshape(D ** 3, D ** 3), W.reshape(D ** 3, D ** 3)
    Mm = vol * np.einsum("il,jm->jiml", A2, A2).reshape(D ** 2, D ** 2)
    K = np.zeros((D,) * 4)
    W = np.zeros((D,) * 4)
    for d in range(2):
        T = [B3 if a == d else A3 for a in range(2)]
        s = vol / h[d] ** 2
        K += s * np.einsum("ba,ail,bjm->jiml", k, T[0], T[1])
        W += s * np.einsum("ba,lai,mbj->jiml", z, T[0], T[1])
    return Mm, K.reshape(D ** 2, D ** 2), W.reshape(D ** 2, D ** 2)


def oracle_operators(dim, p, h, kn, zn, tmp_path=None):
    """the oracle's assembled matrices of a one-element mesh, rows / columns in the element's local (lexicographic) order"""
    if dim == 3:
        mesh = M.make_brick(1, 1, 1, *h)
    else:
        from cases import write_quad_mesh
        path = str(tmp_path / ("q%d.mesh" % p))
        write_quad_mesh(path, 1, 1, h[0], h[1])
        mesh = M.read_mfem_mesh(path)
    s = fem.H1Space(mesh, p)
    g = s.gather[0]
    kg = np.empty(s.ndof)
    zg = np.empty(s.ndof)
    kg[g] = kn
    zg[g] = zn
    X = s.node_coordinates()[g]
    D = p + 1
    assert np.all(np.diff(X[:D, 0]) > 0) and (dim == 2 or X[D * D, 2] > X[0, 2]) and X[D, 1] > X[0, 1]     # x fastest, then y, then z
    pick = np.ix_(g, g)
    return (s.mass_matrix().toarray()[pick], s.diffusion_matrix(kg).toarray()[pick], s.weakdiv_term_matrix(1.0, zg).toarray()[pick])


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("p", [1, 2, 3])
def test_hex_operators_are_the_exact_galerkin_integrals(p):
    rng = np.random.default_rng(100 + p)
    h = (0.3, 0.5, 0.7)
    n = (p + 1) ** 3
    kn = 20.0 + 10.0 * rng.random(n)              # nodal k(T) values
    zn = 300.0 + 200.0 * rng.random(n)            # nodal state
    Mo, Ko, Wo = oracle_operators(3, p, h, kn, zn)
    Me, Ke, We = element_operators(p, h, kn, zn, exact_1d(p))
    assert rel(Mo, Me) < 1e-13
    assert rel(Ko, Ke) < 1e-13
    assert rel(Wo, We) < 1e-13
    # consistency of the exact tensors themselves: K annihilates constants, W z-row sums reproduce K(1) z
    assert np.abs(Ke @ np.ones(n)).max() < 1e-12 * np.abs(Ke).max()
    K1 = element_operators(p, h, np.ones(n), zn, exact_1d(p))[1]
    assert rel(We @ np.ones(n), K1 @ zn) < 1e-12


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_constant_coefficient_operators_are_exact_at_every_order(p, tmp_path):
    """mass and constant-coefficient diffusion have degree 2p: exact for hexes (p + 2 points) and quads (p + 1 points) alike."""
    for dim, h in ((3, (0.3, 0.5, 0.7)), (2, (0.4, 0.9))):
        n = (p + 1) ** dim
        Mo, Ko, _ = oracle_operators(dim, p, h, np.ones(n), np.zeros(n), tmp_path)
        Me, Ke, _ = element_operators(p, h, np.ones(n), np.zeros(n), exact_1d(p))
        assert rel(Mo, Me) < 1e-13, (dim, p)
        assert rel(Ko, Ke) < 1e-13, (dim, p)


@pytest.mark.parametrize("p", [2, 3])
def test_quad_rule_is_the_p_plus_1_point_gauss_sum_not_the_exact_integral(p, tmp_path):
    """Quads: MFEM's default rule has p + 1 points (exact to degree 2p + 1), the k_h-weighted diffusion integrand has degree 3p:
    the oracle must equal the (p + 1)-point Gauss sum and differ from the exact integral -- the quadrature ORDER is part of the
    discrete problem and is pinned here."""
    rng = np.random.default_rng(200 + p)
    h = (0.4, 0.9)
    n = (p + 1) ** 2
    kn = 20.0 + 10.0 * rng.random(n)
    zn = 300.0 + 200.0 * rng.random(n)
    _, Ko, Wo = oracle_operators(2, p, h, kn, zn, tmp_path)
    _, Kg, Wg = element_operators(p, h, kn, zn, gauss_1d(p, p + 1))
    _, Ke, _ = element_operators(p, h, kn, zn, exact_1d(p))
    assert rel(Ko, Kg) < 1e-13
    assert rel(Wo, Wg) < 1e-13
    assert rel(Ko, Ke) > 1e-6
    # one more point would already be a different operator
    _, Kg2, _ = element_operators(p, h, kn, zn, gauss_1d(p, p + 2))
    assert rel(Ko, Kg2) > 1e-6


def test_benchmark_jacobian_on_a_small_brick_is_the_exact_integral(tmp_path):
    """The Newton Jacobian of BASELINE configs[3]'s physics (Nonlinear_Unsteady, k = 0.09 T - 17, uniform rho C, backward Euler)
    as the oracle assembles it on a 2 x 2 x 1 Q2 brick: J = M + dt (K(k_h) + k' W(z)) / (rho C) with essential rows / columns
    eliminated -- against the same matrix assembled here from the exact element integrals."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cases import ini_text
    from oracle import jots as OJ
    p, nx, ny, nz = 2, 2, 2, 1
    h = (0.01 / nx, 0.01 / ny, 0.01 / nz)
    props = [("Density_rho", "Uniform, 8000"), ("Specific_Heat_C", "Uniform, 500"), ("Thermal_Conductivity_k", "Polynomial, 0.09, -17")]
    bcs = ["HeatFlux, 7.5e5", "HeatFlux, 0", "HeatFlux, 0", "Isothermal, 300", "HeatFlux, 0", "HeatFlux, 0"]
    ini = str(tmp_path / "c.ini")
    with open(ini, "w") as f:
        f.write(ini_text("Nonlinear_Unsteady", "unused.mesh", 300, props, bcs, newton=50, solver="CG", prec="Jacobi", steps=1, order=p))
    orc = OJ.JOTSDriver(OJ.Config(ini), mesh=M.make_brick(nx, ny, nz, 0.01, 0.01, 0.01))
    it, s = orc.it, orc.space
    rng = np.random.default_rng(7)
    u_n = 300.0 + 400.0 * rng.random(s.ndof)
    kvec = 1000.0 * (rng.random(s.ndof) - 0.5)
    it.u_n = u_n
    it.new_state(kvec)
    Jo = it.gradient(kvec).toarray()
    z = u_n + it.dt * kvec
    rc = 8000.0 * 500.0
    J = np.zeros((s.ndof, s.ndof))
    for e in range(s.gather.shape[0]):
        g = s.gather[e]
        Me, Ke, We = element_operators(p, h, 0.09 * z[g] - 17.0, z[g], exact_1d(p))
        J[np.ix_(g, g)] += Me + it.dt * (Ke + 0.09 * We) / rc
    ess = np.asarray(it.ess)
    J[ess, :] = 0.0
    J[:, ess] = 0.0
    J[ess, ess] = 1.0
    free = np.setdiff1d(np.arange(s.ndof), ess)
    assert ess.size > 0 and free.size > 0
    # the unit diagonal (1.0) dwarfs the entries of the free block (1e-9): compare the blocks separately
    assert rel(Jo[np.ix_(free, free)], J[np.ix_(free, free)]) < 1e-13
    assert np.array_equal(Jo[ess, :], J[ess, :]) and np.array_equal(Jo[:, ess], J[:, ess])


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_neumann_load_vector_is_the_exact_face_integral(p):
    """BoundaryLFIntegrator of a uniform flux (jots_iterator.cpp:33): b_i = g int_face phi_i dS, integrand of degree p per
    direction, rule with floor((p + 1) / 2) + 1 points (exact to degree >= p): the exact integral, on the x0 and y1 faces of a
    2 x 1 x 2 brick, every other entry zero."""
    h = (0.3, 0.5, 0.7)
    nx, ny, nz = 2, 1, 2
    s = fem.H1Space(M.make_brick(nx, ny, nz, nx * h[0], ny * h[1], nz * h[2]), p)
    g1, g5 = 7.5e5, -2.0e5
    b = s.neumann_vector({1: g1, 5: g5})
    I1 = np.array([float(_pint01(list(f))) for f in basis(p)])          # int phi_i on [0, 1]
    Xn = s.node_coordinates()
    ref = np.zeros(s.ndof)
    for e in range(s.gather.shape[0]):
        ex, ey, ez = e % nx, (e // nx) % ny, e // (nx * ny)
        g = s.gather[e].reshape(p + 1, p + 1, p + 1)                      # [k, j, i]
        if ex == 0:
            ref[g[:, :, 0]] += g1 * h[1] * h[2] * np.outer(I1, I1)        # [k, j]
        if ey == ny - 1:
            ref[g[:, p, :]] += g5 * h[0] * h[2] * np.outer(I1, I1)        # [k, i]
    assert np.abs(Xn[s.gather[0].reshape(p + 1, p + 1, p + 1)[:, :, 0], 0]).max() == 0.0
    assert np.abs(b - ref).max() < 1e-13 * np.abs(ref).max()
