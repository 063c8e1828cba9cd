// Device-side material property and coefficient evaluation shared by the apply / diagonal /
// boundary kernels, so that all of them see bit-identical coefficients.
#pragma once
#include "forms.hpp"

namespace jots {

// PolynomialProperty (/root/reference/src/material_property.cpp:36-81): value, first and second derivative
// of sum_i c_i u^(n-1-i) (coefficients highest power first).  The reference forms every power by repeated
// multiplication (O(n^2) products); here one Horner sweep yields all three (3 FMA per coefficient, no
// per-degree code) -- same polynomial, results equal to within a few ulp.  Uniform (n == 1): value c_0,
// derivatives 0.  n is uniform across the grid, so the loop does not diverge.
__device__ __forceinline__ void poly_eval(const Poly& P, double u, double& val, double& d1, double& d2) {
    double p = P.c[0], dp = 0.0, ddp = 0.0;
#pragma unroll 1
    for (int i = 1; i < P.n; ++i) {
        ddp = ddp * u + dp;
        dp = dp * u + p;
        p = p * u + P.c[i];
    }
    val = p; d1 = dp; d2 = 2.0 * ddp;
}
__device__ __forceinline__ double poly_val(const Poly& P, double u) {
    double p = P.c[0];
#pragma unroll 1
    for (int i = 1; i < P.n; ++i) p = p * u + P.c[i];
    return p;
}
__device__ __forceinline__ double poly_d1(const Poly& P, double u) {
    double v, d1, d2;
    poly_eval(P, u, v, d1, d2);
    return d1;
}
__device__ __forceinline__ double poly_d2(const Poly& P, double u) {
    double v, d1, d2;
    poly_eval(P, u, v, d1, d2);
    return d2;
}

template <int CF, int OP> struct CoefNF { static constexpr int value = 0; };
template <int OP> struct CoefNF<CF_LINFIELDS, OP> { static constexpr int value = 3; };
template <> struct CoefNF<CF_K, OP_LIN> { static constexpr int value = 1; };
template <> struct CoefNF<CF_K, OP_RES> { static constexpr int value = 1; };
template <> struct CoefNF<CF_K, OP_JAC> { static constexpr int value = 2; };
template <int OP> struct CoefNF<CF_KL, OP> { static constexpr int value = 1; };
template <int OP> struct CoefNF<CF_FULL, OP> { static constexpr int value = 8; };
template <int OP> struct CoefNF<CF_FULLC, OP> { static constexpr int value = 5; };

// nodal fields to interpolate, from the nodal state value u
template <int CF, int OP>
__device__ __forceinline__ void nodal_fields(const MatSet& mat, double u, double* f) {
    double v, d1, d2;
    if (CF == CF_LINFIELDS) {
        f[0] = poly_val(mat.rho, u); f[1] = poly_val(mat.C, u); f[2] = poly_val(mat.k, u);
    } else if (CF == CF_KL) {
        f[0] = poly_val(mat.k, u);
    } else if (CF == CF_K) {
        if (OP == OP_JAC) { poly_eval(mat.k, u, v, d1, d2); f[0] = v; f[1] = d1; }
        else f[0] = poly_val(mat.k, u);
    } else if (CF == CF_FULLC) {
        poly_eval(mat.k, u, v, d1, d2); f[0] = v; f[1] = d1;
        poly_eval(mat.C, u, v, d1, d2); f[2] = v; f[3] = d1; f[4] = d2;
    } else if (CF == CF_FULL) {
        poly_eval(mat.k, u, v, d1, d2); f[0] = v; f[1] = d1;
        poly_eval(mat.rho, u, v, d1, d2); f[2] = v; f[3] = d1; f[4] = d2;
        poly_eval(mat.C, u, v, d1, d2); f[5] = v; f[6] = d1; f[7] = d2;
    }
}

struct QCoef {
    double m, a1, a2, nb, ndb;
};

// quadrature-point coefficients from the interpolated nodal fields
// (DiffusivityCoefficient ... dBetaCoefficient::Eval, nl_conduction_operator.cpp:92-163)
template <int CF, int OP>
__device__ __forceinline__ QCoef qpoint_coef(const FormArgs& a, const double* f) {
    QCoef c;
    c.m = a.cm; c.a1 = 0.0; c.a2 = 0.0; c.nb = 0.0; c.ndb = 0.0;
    if (CF == CF_CONST) {
        c.m = a.cm * a.m0; c.a1 = a.ck * a.a0;
    } else if (CF == CF_LINFIELDS) {
        c.m = a.cm * (f[0] * f[1]); c.a1 = a.ck * f[2];
    } else if (CF == CF_K) {
        c.m = a.cm * a.m0;
        c.a1 = a.ck * (f[0] * a.inv_rc);
        if (OP == OP_JAC) c.a2 = a.ck * (f[1] * a.inv_rc);
    } else {
        const double k = f[0], dk = f[1], r = f[2], dr = f[3], d2r = f[4], cc = f[5], dc = f[6], d2c = f[7];
        const double alpha = k / (r * cc);
        c.a1 = a.ck * alpha;
        const double t = -dr / (r * r * cc) - dc / (cc * cc * r);
        c.nb = a.cb * (-(k * t));
        if (OP == OP_JAC) {
            c.a2 = a.ck * (dk / (r * cc) - k * dr / (r * r * cc) - k * dc / (cc * cc * r));
            c.ndb = a.cb * (-(dk * t + k * (-d2c / (cc * cc * r) + 2.0 * dc * dr / (cc * cc * r * r) +
                                             2.0 * dc * dc / (cc * cc * cc * r) - d2r / (cc * r * r) +
                                             2.0 * dr * dr / (cc * r * r * r))));
        }
    }
    return c;
}

}  // namespace jots
