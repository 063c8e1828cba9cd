// Loop-based (not unrolled) kernels that are off the per-iteration hot path:
//   * diagonal of an OP_LIN / OP_JAC form (replaces hypre's diagonal extraction inside
//     HypreSmoother::SetOperator; called once per operator change / Newton iteration);
//   * boundary-face kernels: Neumann load vector (BoundaryLFIntegrator,
//     /root/reference/src/jots_iterator.cpp:33,39-43), the g/(rho C) variants of
//     JOTSNonlinearNeumannIntegrator (jots_nlfis.cpp:42-59) and its boundary-mass Jacobian.
#pragma once
#include "coef.cuh"
#include "diag_face_host.hpp"

namespace jots {

// one block per element (grid-stride); blockDim.x threads
template <int OP, int CF>
__global__ void __launch_bounds__(128) form_diag_kernel(const FormArgs a, const TablesRT t) {
    constexpr int NF = CoefNF<CF, OP>::value;
    constexpr int NFA = NF > 0 ? NF : 1;
    const int D = t.D, Q = t.Q, DZ = t.DZ, QZ = t.QZ;
    const int ND = D * D * DZ, NQ = Q * Q * QZ;
    extern __shared__ double smem[];
    double* zn = smem;                 // [ND]
    double* fn = zn + ND;              // [NFA][ND]
    double* qd = fn + NFA * ND;        // [11][NQ]: wm, wa1, wr0, wr1, wr2, metric g00 g01 g02 g11 g12 g22
    for (int e = blockIdx.x; e < a.nelem; e += gridDim.x) {
        __syncthreads();
        for (int l = threadIdx.x; l < ND; l += blockDim.x) {
            const int gg = a.gather[(size_t)e * ND + l];
            const int idx = gg < 0 ? ~gg : gg;
            double zv = 0.0;
            if (OP != OP_LIN || NF > 0) zv = a.z[idx];
            zn[l] = zv;
            if (NF > 0) {
                double f[NFA];
                nodal_fields<CF, OP>(a.mat, zv, f);
                for (int n = 0; n < NF; ++n) fn[n * ND + l] = f[n];
            }
        }
        __syncthreads();
        for (int q = threadIdx.x; q < NQ; q += blockDim.x) {
            const int q0 = q % Q, q1 = (q / Q) % Q, q2 = q / (Q * Q);
            const double* g = a.geomq ? a.geomq + ((size_t)e * NQ + q) * 10 : a.geom + (size_t)e * a.geom_stride;
            double iJ[9];
            for (int i = 0; i < 9; ++i) iJ[i] = g[i];
            const double detJ = g[9];
            double f[NFA];
            for (int n = 0; n < NFA; ++n) f[n] = 0.0;
            double gr[3] = {0, 0, 0};
            for (int l = 0; l < ND; ++l) {
                const int i0 = l % D, i1 = (l / D) % D, i2 = l / (D * D);
                const double b0 = t.B[q0 * D + i0], b1 = t.B[q1 * D + i1], b2 = t.Bz[q2 * DZ + i2];
                const double phi = b0 * b1 * b2;
                for (int n = 0; n < NF; ++n) f[n] += phi * fn[n * ND + l];
                if (OP != OP_LIN) {
                    const double zv = zn[l];
                    gr[0] += t.G[q0 * D + i0] * b1 * b2 * zv;
                    gr[1] += b0 * t.G[q1 * D + i1] * b2 * zv;
                    gr[2] += b0 * b1 * t.Gz[q2 * DZ + i2] * zv;
                }
            }
            const QCoef c = qpoint_coef<CF, OP>(a, f);
            const double w = t.W[q0] * t.W[q1] * t.Wz[q2] * detJ;
            double pz[3];
            for (int i = 0; i < 3; ++i) pz[i] = iJ[0 + i] * gr[0] + iJ[3 + i] * gr[1] + iJ[6 + i] * gr[2];
            const double zz = pz[0] * pz[0] + pz[1] * pz[1] + pz[2] * pz[2];
            double wm = w * c.m, cr = 0.0;
            if (OP == OP_JAC) { wm += w * c.ndb * zz; cr = w * (c.a2 + 2.0 * c.nb); }
            qd[0 * NQ + q] = wm;
            qd[1 * NQ + q] = w * c.a1;
            for (int aa = 0; aa < 3; ++aa)
                qd[(2 + aa) * NQ + q] = cr * (iJ[3 * aa + 0] * pz[0] + iJ[3 * aa + 1] * pz[1] + iJ[3 * aa + 2] * pz[2]);
            int n = 5;
            for (int aa = 0; aa < 3; ++aa)
                for (int bb = aa; bb < 3; ++bb)
                    qd[(n++) * NQ + q] = iJ[3 * aa] * iJ[3 * bb] + iJ[3 * aa + 1] * iJ[3 * bb + 1] + iJ[3 * aa + 2] * iJ[3 * bb + 2];
        }
        __syncthreads();
        for (int l = threadIdx.x; l < ND; l += blockDim.x) {
            const int gg = a.gather[(size_t)e * ND + l];
            if (gg < 0 && a.mask_out) continue;
            const int i0 = l % D, i1 = (l / D) % D, i2 = l / (D * D);
            double d = 0.0;
            for (int q = 0; q < NQ; ++q) {
                const int q0 = q % Q, q1 = (q / Q) % Q, q2 = q / (Q * Q);
                const double b0 = t.B[q0 * D + i0], b1 = t.B[q1 * D + i1], b2 = t.Bz[q2 * DZ + i2];
                const double phi = b0 * b1 * b2;
                const double d0 = t.G[q0 * D + i0] * b1 * b2, d1 = b0 * t.G[q1 * D + i1] * b2, d2 = b0 * b1 * t.Gz[q2 * DZ + i2];
                const double gg2 = qd[5 * NQ + q] * d0 * d0 + qd[8 * NQ + q] * d1 * d1 + qd[10 * NQ + q] * d2 * d2 +
                                   2.0 * (qd[6 * NQ + q] * d0 * d1 + qd[7 * NQ + q] * d0 * d2 + qd[9 * NQ + q] * d1 * d2);
                d += qd[1 * NQ + q] * gg2 + qd[0 * NQ + q] * phi * phi;
                if (OP == OP_JAC) d += phi * (qd[2 * NQ + q] * d0 + qd[3 * NQ + q] * d1 + qd[4 * NQ + q] * d2);
            }
            atomicAdd(a.y + (gg < 0 ? ~gg : gg), d);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// boundary faces
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void face_kernel(const FaceArgs a) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.nbdr) return;
    const int attr = a.bdr_attr[b];
    const bool mapped = attr >= 1 && attr <= a.n_attr;
    const double* gfield = (mapped && a.g_nodal) ? a.g_nodal[attr - 1] : nullptr;   // nodal flux field of this attribute
    const double g0 = mapped ? a.g_attr[attr - 1] : 0.0;
    if (g0 == 0.0 && !gfield) return;
    const int dim = a.dim, D = a.D, nq = a.nq, nfl = a.nfl;
    const int nq2 = dim == 3 ? nq : 1;
    const int* gt = a.bdr_gather + (size_t)b * nfl;
    const double* X = a.bdr_corner + (size_t)b * (dim == 3 ? 4 : 2) * dim;
    double rho_n[25], C_n[25], dr_n[25], dC_n[25], xn[25], acc[25], gn[25];
    const bool state = a.z != nullptr;
    for (int l = 0; l < nfl; ++l) {
        acc[l] = 0.0;
        const int idx = gt[l];
        gn[l] = gfield ? gfield[idx] : g0;
        if (state) {
            const double zv = a.z[idx];
            rho_n[l] = poly_val(a.mat.rho, zv); C_n[l] = poly_val(a.mat.C, zv);
            dr_n[l] = poly_d1(a.mat.rho, zv); dC_n[l] = poly_d1(a.mat.C, zv);
        }
        if (MODE == FACE_MASS_APPLY) xn[l] = (a.mask_in && a.ess && a.ess[idx]) ? 0.0 : a.x[idx];
    }
    for (int qb = 0; qb < nq2; ++qb)
        for (int qa = 0; qa < nq; ++qa) {
            // surface Jacobian of the (multi)linear face at (s,t)
            double meas;
            if (dim == 2) {
                const double dx = X[2] - X[0], dy = X[3] - X[1];
                meas = sqrt(dx * dx + dy * dy);
            } else {
                const double s = a.xq[qa], tt = a.xq[qb];
                double ts[3], tv[3];
                for (int i = 0; i < 3; ++i) {
                    ts[i] = (1.0 - tt) * (X[3 + i] - X[i]) + tt * (X[9 + i] - X[6 + i]);
                    tv[i] = (1.0 - s) * (X[6 + i] - X[i]) + s * (X[9 + i] - X[3 + i]);
                }
                const double c0 = ts[1] * tv[2] - ts[2] * tv[1], c1 = ts[2] * tv[0] - ts[0] * tv[2], c2 = ts[0] * tv[1] - ts[1] * tv[0];
                meas = sqrt(c0 * c0 + c1 * c1 + c2 * c2);
            }
            const double w = a.Wf[qa] * (dim == 3 ? a.Wf[qb] : 1.0) * meas;
            double r = 0.0, c = 0.0, dr = 0.0, dc = 0.0, xv = 0.0, g = gfield ? 0.0 : g0;
            for (int l = 0; l < nfl; ++l) {
                const int la = l % D, lb = l / D;
                const double phi = a.Bf[qa * D + la] * (dim == 3 ? a.Bf[qb * D + lb] : 1.0);
                if (gfield) g += phi * gn[l];          // GridFunctionCoefficient: the nodal field interpolated on the face
                if (state) { r += phi * rho_n[l]; c += phi * C_n[l]; dr += phi * dr_n[l]; dc += phi * dC_n[l]; }
                if (MODE == FACE_MASS_APPLY) xv += phi * xn[l];
            }
            double lam;
            if (MODE == FACE_LOAD) lam = state ? g / (r * c) : g / a.div_const;
            else lam = -g * dr / (r * r * c) - g * dc / (c * c * r);   // d(g/(rho C))/du
            for (int l = 0; l < nfl; ++l) {
                const int la = l % D, lb = l / D;
                const double phi = a.Bf[qa * D + la] * (dim == 3 ? a.Bf[qb * D + lb] : 1.0);
                if (MODE == FACE_LOAD) acc[l] += w * lam * phi;
                else if (MODE == FACE_MASS_APPLY) acc[l] += w * lam * phi * xv;
                else acc[l] += w * lam * phi * phi;
            }
        }
    for (int l = 0; l < nfl; ++l) {
        const int idx = gt[l];
        if (a.mask_out && a.ess && a.ess[idx]) continue;
        atomicAdd(a.y + idx, a.scale * acc[l]);
    }
}

// one thread per (boundary element of the BC, face node): gradient of the adjacent element's field at that node through the
// nodal derivative matrix, outward normal = +- grad xi_axis
static __global__ void wall_flux_kernel(const WallFluxArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nfaces * a.nfl) return;
    const int fi = t / a.nfl, l = t - fi * a.nfl;
    const int b = a.faces[fi], e = a.bdr_elem[b], lf = a.bdr_face[b];
    const int dim = a.dim, D = a.D, axis = lf >> 1, side = lf & 1;
    // the node's place in the adjacent element comes from a table (the boundary element may list its DOFs in any order)
    const int l0 = a.node_loc[(size_t)b * a.nfl + l];
    const int idx[3] = {l0 % D, (l0 / D) % D, l0 / (D * D)};
    const int stride[3] = {1, D, D * D};
    const int* g = a.gather + (size_t)e * a.nloc;
    auto val = [&](int loc) { const int d = g[loc]; return a.u[d < 0 ? ~d : d]; };
    double gref[3] = {0.0, 0.0, 0.0};
    for (int c = 0; c < dim; ++c) {
        const int base = l0 - idx[c] * stride[c];
        double s = 0.0;
        for (int m = 0; m < D; ++m) s += a.Dn[idx[c] * D + m] * val(base + m * stride[c]);
        gref[c] = s;
    }
    const double* iJ = a.geom + (size_t)e * a.geom_stride;      // iJ[3*a + i] = d xi_a / d x_i
    double n[3], nn = 0.0, flux = 0.0;
    for (int i = 0; i < 3; ++i) { n[i] = (side ? 1.0 : -1.0) * iJ[3 * axis + i]; nn += n[i] * n[i]; }
    nn = sqrt(nn);
    for (int i = 0; i < dim; ++i) {
        double gp = 0.0;
        for (int c = 0; c < dim; ++c) gp += iJ[3 * c + i] * gref[c];
        flux += gp * n[i];
    }
    a.out[t] = -poly_val(a.k, val(l0)) * flux / nn;
}

}  // namespace jots
