// Host-side 1-D rules and Lagrange tables for the H1 tensor-product space.
//
// What the reference gets from MFEM (not vendored in /root/reference):
//   * H1_FECollection(p, dim): nodal Lagrange basis on Gauss-Lobatto points
//     (/root/reference/src/jots_driver.cpp:194);
//   * default integration rules of Mass/Diffusion/Convection/MixedScalarWeakDivergence
//     integrators on d-linear tensor elements: Gauss-Legendre with p+2 points per direction on
//     hexes, p+1 on quads; BoundaryLFIntegrator: (p+1)/2+1 points; boundary MassIntegrator: p+1.
#pragma once
#include <cmath>
#include <vector>

namespace jots {

inline void legendre(int n, long double x, long double& P, long double& dP) {
    // P_n(x) and P_n'(x) on [-1,1]
    long double p0 = 1.0L, p1 = x;
    if (n == 0) { P = 1.0L; dP = 0.0L; return; }
    for (int k = 2; k <= n; ++k) {
        long double pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
        p0 = p1; p1 = pk;
    }
    P = p1;
    dP = n * (x * p1 - p0) / (x * x - 1.0L);
}

// n Gauss-Legendre points / weights on [0,1], ascending
inline void gauss_legendre_01(int n, std::vector<double>& x, std::vector<double>& w) {
    x.assign(n, 0.0); w.assign(n, 0.0);
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int i = 0; i < (n + 1) / 2; ++i) {
        long double z = cosl(pi * (i + 0.75L) / (n + 0.5L));
        long double P, dP;
        for (int it = 0; it < 100; ++it) {
            legendre(n, z, P, dP);
            long double dz = P / dP;
            z -= dz;
            if (fabsl(dz) < 1e-19L) break;
        }
        legendre(n, z, P, dP);
        long double wt = 2.0L / ((1.0L - z * z) * dP * dP);
        // z is the i-th largest root
        x[n - 1 - i] = (double)(0.5L * (1.0L + z));
        x[i] = (double)(0.5L * (1.0L - z));
        w[n - 1 - i] = w[i] = (double)(0.5L * wt);
    }
    if (n % 2 == 1) x[n / 2] = 0.5;
}

// n Gauss-Lobatto points on [0,1] (n >= 2), ascending
inline std::vector<double> gauss_lobatto_01(int n) {
    std::vector<double> x(n);
    x[0] = 0.0; x[n - 1] = 1.0;
    const long double pi = 3.141592653589793238462643383279502884L;
    const int m = n - 1;  // interior points are the roots of P_m'
    for (int i = 1; i < n - 1; ++i) {
        if (2 * i == m) { x[i] = 0.5; continue; }
        long double z = -cosl(pi * i / m);
        for (int it = 0; it < 100; ++it) {
            long double P, dP;
            legendre(m, z, P, dP);
            // d2P from the Legendre ODE: (1-z^2) P'' - 2 z P' + m(m+1) P = 0
            long double d2P = (2 * z * dP - m * (m + 1) * P) / (1.0L - z * z);
            long double dz = dP / d2P;
            z -= dz;
            if (fabsl(dz) < 1e-19L) break;
        }
        x[i] = (double)(0.5L * (1.0L + z));
    }
    // symmetrise
    for (int i = 0; i < n / 2; ++i) {
        double a = 0.5 * (x[i] + (1.0 - x[n - 1 - i]));
        x[i] = a; x[n - 1 - i] = 1.0 - a;
    }
    return x;
}

// B[q*n+d] = l_d(x_q), G[q*n+d] = l_d'(x_q) for the Lagrange basis on `nodes`
inline void lagrange_tab(const std::vector<double>& nodes, const std::vector<double>& x,
                         std::vector<double>& B, std::vector<double>& G) {
    const int n = (int)nodes.size(), nq = (int)x.size();
    B.assign((size_t)nq * n, 0.0); G.assign((size_t)nq * n, 0.0);
    for (int q = 0; q < nq; ++q)
        for (int d = 0; d < n; ++d) {
            long double den = 1.0L;
            for (int m = 0; m < n; ++m) if (m != d) den *= (long double)nodes[d] - nodes[m];
            long double b = 1.0L, g = 0.0L;
            for (int m = 0; m < n; ++m) {
                if (m == d) continue;
                b *= (long double)x[q] - nodes[m];
                long double term = 1.0L;
                for (int r = 0; r < n; ++r) if (r != d && r != m) term *= (long double)x[q] - nodes[r];
                g += term;
            }
            B[(size_t)q * n + d] = (double)(b / den);
            G[(size_t)q * n + d] = (double)(g / den);
        }
}

inline int domain_quad_points(int dim, int p) { return dim == 3 ? p + 2 : p + 1; }
inline int bdr_lf_quad_points(int p) { return (p + 1) / 2 + 1; }
inline int bdr_mass_quad_points(int p) { return p + 1; }

}  // namespace jots
