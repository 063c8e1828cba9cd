// Device-side material property and coefficient evaluation shared by the apply / diagonal /
// boundary kernels, so that all of them see bit-identical coefficients.
#pragma once
#include "forms.hpp"

namespace jots {

// PolynomialProperty::UpdateCoeff / UpdateDCoeff / UpdateD2Coeff arithmetic
// (/root/reference/src/material_property.cpp:36-81): sum_i c_i u^(n-1-i), powers by repeated
// multiplication, highest power first.  Uniform (n == 1): value c_0, derivatives 0.
__device__ __forceinline__ double poly_val(const Poly& P, double u) {
    if (P.n == 1) return P.c[0];
    double z1 = 0.0;
    for (int i = 0; i < P.n; ++i) {
        double z2 = 1.0;
        for (int j = 0; j < P.n - i - 1; ++j) z2 *= u;
        z1 += z2 * P.c[i];
    }
    return z1;
}
__device__ __forceinline__ double poly_d1(const Poly& P, double u) {
    if (P.n == 1) return 0.0;
    double z1 = 0.0;
    for (int i = 0; i < P.n - 1; ++i) {
        double z2 = (double)(P.n - 1 - i);
        for (int j = 0; j < P.n - i - 2; ++j) z2 *= u;
        z1 += z2 * P.c[i];
    }
    return z1;
}
__device__ __forceinline__ double poly_d2(const Poly& P, double u) {
    if (P.n == 1) return 0.0;
    double z1 = 0.0;
    for (int i = 0; i < P.n - 2; ++i) {
        double z2 = (double)((P.n - 1 - i) * (P.n - 2 - i));
        for (int j = 0; j < P.n - i - 3; ++j) z2 *= u;
        z1 += z2 * P.c[i];
    }
    return z1;
}

template <int CF, int OP> struct CoefNF { static constexpr int value = 0; };
template <int OP> struct CoefNF<CF_LINFIELDS, OP> { static constexpr int value = 3; };
template <> struct CoefNF<CF_K, OP_LIN> { static constexpr int value = 1; };
template <> struct CoefNF<CF_K, OP_RES> { static constexpr int value = 1; };
template <> struct CoefNF<CF_K, OP_JAC> { static constexpr int value = 2; };
template <int OP> struct CoefNF<CF_FULL, OP> { static constexpr int value = 8; };
template <int OP> struct CoefNF<CF_FULLC, OP> { static constexpr int value = 5; };

// nodal fields to interpolate, from the nodal state value u
template <int CF, int OP>
__device__ __forceinline__ void nodal_fields(const MatSet& mat, double u, double* f) {
    if (CF == CF_LINFIELDS) {
        f[0] = poly_val(mat.rho, u); f[1] = poly_val(mat.C, u); f[2] = poly_val(mat.k, u);
    } else if (CF == CF_K) {
        f[0] = poly_val(mat.k, u);
        if (OP == OP_JAC) f[1] = poly_d1(mat.k, u);
    } else if (CF == CF_FULLC) {
        f[0] = poly_val(mat.k, u); f[1] = poly_d1(mat.k, u);
        f[2] = poly_val(mat.C, u); f[3] = poly_d1(mat.C, u); f[4] = poly_d2(mat.C, u);
    } else if (CF == CF_FULL) {
        f[0] = poly_val(mat.k, u); f[1] = poly_d1(mat.k, u);
        f[2] = poly_val(mat.rho, u); f[3] = poly_d1(mat.rho, u); f[4] = poly_d2(mat.rho, u);
        f[5] = poly_val(mat.C, u); f[6] = poly_d1(mat.C, u); f[7] = poly_d2(mat.C, u);
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
