"""Oracle (test infrastructure): the analytic solutions and the error functional the
reference's regression tests use (/root/reference/tests/test_helper.hpp.in:107-159 and
:261-376).  These are the only result pins the reference ships."""
from __future__ import annotations

import numpy as np

N_REINERT = 1000  # test_helper.hpp.in:14


def reinert_b1(x, t):
    """test_helper.hpp.in:261-281"""
    xx = x[:, 0]
    alpha, L = 2.5e-6, 0.01
    n = np.arange(N_REINERT + 1)[:, None]
    A = (-1.0) ** n / (2.0 * n + 1.0)
    B = (-(2.0 * n + 1.0) ** 2 * np.pi ** 2 * alpha * t) / (4.0 * L ** 2)
    C = ((2.0 * n + 1.0) * np.pi * xx[None, :]) / (2.0 * L)
    theta = 1.0 - np.sum((4.0 / np.pi) * A * np.exp(B) * np.cos(C), axis=0)
    return theta * (500.0 - 300.0) + 300.0


def _b2_theta(xx, t):
    alpha, L = 2.5e-6, 0.01
    theta = alpha * t / L ** 2 + 1.0 / 3.0 - xx / L + 0.5 * (xx / L) ** 2
    n = np.arange(1, N_REINERT + 1)[:, None]
    A = -n ** 2 * np.pi ** 2 * alpha * t / L ** 2
    B = n * np.pi * xx[None, :] / L
    C = (2.0 / np.pi ** 2) * (1.0 / n ** 2)
    return theta - np.sum(C * np.exp(A) * np.cos(B), axis=0)


def reinert_b2(x, t):
    """test_helper.hpp.in:307-328"""
    L, k, q = 0.01, 10.0, 7.5e5
    return _b2_theta(x[:, 0], t) * q * L / k + 300.0


def reinert_b3(x, t):
    """test_helper.hpp.in:330-364 (Kirchhoff transform)"""
    L, k1, k2, T1, T2, q, T0 = 0.01, 10.0, 100.0, 300.0, 1300.0, 7.5e5, 300.0
    D = ((k2 - k1) / (T2 - T1)) * (1.0 / (2.0 * k1))
    theta0 = (T0 - T1) + D * (T0 - T1) ** 2
    theta = _b2_theta(x[:, 0], t) * q * L / k1 + theta0
    a = D
    b = 1.0 - 2.0 * D * T1
    c = D * T1 ** 2 - T1 - theta
    return (-b + np.sqrt(b * b - 4 * a * c)) / (2 * a)


def danish_model1(x, t=0.0):
    """test_helper.hpp.in:366-376"""
    beta = 50.0
    xx = x[:, 0]
    return (-1.0 + np.sqrt(1.0 + 2.0 * beta * xx + beta ** 2 * xx)) / beta


def reference_error(space, u, exact, t):
    """``u_h.ComputeL2Error(exact) / GlobalLpNorm(2, u_exact_nodal.Norml2())``
    (test_helper.hpp.in:134): L2(Omega) norm of the error over the l2 norm of the nodal
    projection of the exact solution."""
    fn = lambda pts: exact(pts, t)
    num = space.l2_error(u, fn)
    den = np.linalg.norm(fn(space.node_coordinates()))
    return num / den
