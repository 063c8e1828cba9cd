// Generic sum-factorised element kernel (all orders Q1-Q4, hexes and quads, every form).
//
// Thread block = (Q, Q, NE): thread (tx, ty) of element slot tz owns the column of quadrature
// points (tx, ty, :) in registers; x/y contractions go through two shared-memory slabs per
// element, z contractions stay in registers with the 1-D tables read from the kernel-parameter
// constant bank.  Gradients use the collocated derivative at the quadrature points.
// L-vector DOFs are gathered through an int32 table and scattered back with FP64 atomics
// (red.global.add.f64).
#pragma once
#include "coef.cuh"

namespace jots {

template <int D, int Q, int DZ, int QZ>
struct ThreadTabs {
    double Bx[D], By[D];      // B[tx][i], B[ty][j]
    double BTx[Q], BTy[Q];    // B[q][tx], B[q][ty]   (tx, ty < D)
    double Gx[Q], Gy[Q];      // Gq[tx][q'], Gq[ty][q']
    double GTx[Q], GTy[Q];    // Gq[q'][tx], Gq[q'][ty]
};

template <int D, int Q, int DZ, int QZ>
__device__ __forceinline__ void load_thread_tabs(const Tables<D, Q, DZ, QZ>& t, int tx, int ty, ThreadTabs<D, Q, DZ, QZ>& r) {
#pragma unroll
    for (int i = 0; i < D; ++i) { r.Bx[i] = t.B[tx * D + i]; r.By[i] = t.B[ty * D + i]; }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        r.BTx[q] = tx < D ? t.B[q * D + tx] : 0.0;
        r.BTy[q] = ty < D ? t.B[q * D + ty] : 0.0;
        r.Gx[q] = t.Gq[tx * Q + q]; r.Gy[q] = t.Gq[ty * Q + q];
        r.GTx[q] = t.Gq[q * Q + tx]; r.GTy[q] = t.Gq[q * Q + ty];
    }
}

// nodal column r[DZ] (threads tx,ty < D) -> quadrature column out[QZ] (all threads)
template <int D, int Q, int DZ, int QZ>
__device__ __forceinline__ void interp_field(const double (&r)[DZ], double (&out)[QZ], double* sA, double* sB,
                                             const Tables<D, Q, DZ, QZ>& t, const ThreadTabs<D, Q, DZ, QZ>& tt,
                                             int tx, int ty) {
    __syncthreads();
    if (tx < D && ty < D) {
#pragma unroll
        for (int qz = 0; qz < QZ; ++qz) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < DZ; ++k) v += t.Bz[qz * DZ + k] * r[k];
            sA[(qz * D + ty) * D + tx] = v;
        }
    }
    __syncthreads();
    if (ty < D) {
#pragma unroll
        for (int qz = 0; qz < QZ; ++qz) {
            double v = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) v += tt.Bx[i] * sA[(qz * D + ty) * D + i];
            sB[(qz * D + ty) * Q + tx] = v;
        }
    }
    __syncthreads();
#pragma unroll
    for (int qz = 0; qz < QZ; ++qz) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) v += tt.By[j] * sB[(qz * D + j) * Q + tx];
        out[qz] = v;
    }
}

// reference gradient of a quadrature column
template <int D, int Q, int DZ, int QZ>
__device__ __forceinline__ void grad_field(const double (&u)[QZ], double (&g0)[QZ], double (&g1)[QZ], double (&g2)[QZ],
                                           double* sA, const Tables<D, Q, DZ, QZ>& t,
                                           const ThreadTabs<D, Q, DZ, QZ>& tt, int tx, int ty) {
    __syncthreads();
#pragma unroll
    for (int qz = 0; qz < QZ; ++qz) {
        sA[(qz * Q + ty) * Q + tx] = u[qz];
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < QZ; ++q) v += t.Gqz[qz * QZ + q] * u[q];
        g2[qz] = v;
    }
    __syncthreads();
#pragma unroll
    for (int qz = 0; qz < QZ; ++qz) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            a += tt.Gx[q] * sA[(qz * Q + ty) * Q + q];
            b += tt.Gy[q] * sA[(qz * Q + q) * Q + tx];
        }
        g0[qz] = a; g1[qz] = b;
    }
}

template <int D, int Q, int DZ, int QZ, int OP, int CF, int NE>
__global__ void __launch_bounds__(Q* Q* NE) form_apply_kernel(const FormArgs a, const Tables<D, Q, DZ, QZ> t) {
    if (a.skip && *a.skip != 0.0) return;   // enqueued ahead of a Krylov solve that has already stopped
    constexpr int NF = CoefNF<CF, OP>::value;
    constexpr int NFA = NF > 0 ? NF : 1;
    constexpr int ND = D * D * DZ;
    constexpr bool NEED_Z = (OP != OP_LIN);             // grad z
    constexpr bool NEED_GX = (OP != OP_RES);            // grad x
    extern __shared__ double smem[];
    const int tx = threadIdx.x, ty = threadIdx.y, tz = threadIdx.z;
    double* sA = smem + tz * (2 * QZ * Q * Q);
    double* sB = sA + QZ * Q * Q;
    int e = blockIdx.x * NE + tz;
    const bool active = e < a.nelem;
    if (!active) e = a.nelem - 1;
    const bool node = (tx < D) && (ty < D);

    ThreadTabs<D, Q, DZ, QZ> tt;
    load_thread_tabs(t, tx, ty, tt);

    // geometry
    double iJ[9], detJ;
    {
        const double* g = a.geom + (size_t)e * a.geom_stride;
#pragma unroll
        for (int i = 0; i < 9; ++i) iJ[i] = g[i];
        detJ = g[9];
    }

    // gather
    double xr[DZ], zr[DZ];
    int gi[DZ];
    double fr[NFA][DZ];
#pragma unroll
    for (int k = 0; k < DZ; ++k) { xr[k] = 0.0; zr[k] = 0.0; gi[k] = 0; }
    if (node) {
#pragma unroll
        for (int k = 0; k < DZ; ++k) {
            const int g = a.gather[(size_t)e * ND + tx + D * (ty + D * k)];
            gi[k] = g;
            const int idx = g < 0 ? ~g : g;
            xr[k] = (g < 0 && a.mask_in) ? 0.0 : a.x[idx];
            if (NEED_Z || NF > 0) zr[k] = a.z[idx];
            if (NF > 0) {
                double f[NFA];
                nodal_fields<CF, OP>(a.mat, a.zc ? a.zc[idx] : zr[k], f);
#pragma unroll
                for (int n = 0; n < NF; ++n) fr[n][k] = f[n];
            }
        }
    }

    // to quadrature points
    double xq[QZ];
    interp_field<D, Q, DZ, QZ>(xr, xq, sA, sB, t, tt, tx, ty);
    double fq[NFA][QZ];
#pragma unroll
    for (int n = 0; n < NF; ++n) interp_field<D, Q, DZ, QZ>(fr[n], fq[n], sA, sB, t, tt, tx, ty);
    double gx0[QZ], gx1[QZ], gx2[QZ];
    if (NEED_GX) grad_field<D, Q, DZ, QZ>(xq, gx0, gx1, gx2, sA, t, tt, tx, ty);
    double gz0[QZ], gz1[QZ], gz2[QZ];
    if (NEED_Z) {
        double zq[QZ];
        interp_field<D, Q, DZ, QZ>(zr, zq, sA, sB, t, tt, tx, ty);
        grad_field<D, Q, DZ, QZ>(zq, gz0, gz1, gz2, sB, t, tt, tx, ty);
    }

    // quadrature-point physics; F (reference components, weighted) overwrites gx*, S -> v
    double v[QZ], F0[QZ], F1[QZ], F2[QZ];
#pragma unroll
    for (int qz = 0; qz < QZ; ++qz) {
        double f[NFA];
#pragma unroll
        for (int n = 0; n < NF; ++n) f[n] = fq[n][qz];
        const QCoef c = qpoint_coef<CF, OP>(a, f);
        if (a.geomq) {
            // general multilinear element: geometric factors of this quadrature point
            const double* g = a.geomq + ((size_t)e * (Q * Q * QZ) + tx + Q * (ty + Q * qz)) * 10;
#pragma unroll
            for (int i = 0; i < 9; ++i) iJ[i] = g[i];
            detJ = g[9];
        }
        const double w = t.W[tx] * t.W[ty] * t.Wz[qz] * detJ;
        double px[3] = {0, 0, 0}, pz[3] = {0, 0, 0};
        if (NEED_GX) {
#pragma unroll
            for (int i = 0; i < 3; ++i) px[i] = iJ[0 + i] * gx0[qz] + iJ[3 + i] * gx1[qz] + iJ[6 + i] * gx2[qz];
        }
        if (NEED_Z) {
#pragma unroll
            for (int i = 0; i < 3; ++i) pz[i] = iJ[0 + i] * gz0[qz] + iJ[3 + i] * gz1[qz] + iJ[6 + i] * gz2[qz];
        }
        double Fp[3], S;
        if (OP == OP_LIN) {
#pragma unroll
            for (int i = 0; i < 3; ++i) Fp[i] = c.a1 * px[i];
            S = c.m * xq[qz];
        } else if (OP == OP_JAC) {
            const double zz = pz[0] * pz[0] + pz[1] * pz[1] + pz[2] * pz[2];
            const double zx = pz[0] * px[0] + pz[1] * px[1] + pz[2] * px[2];
#pragma unroll
            for (int i = 0; i < 3; ++i) Fp[i] = c.a1 * px[i] + c.a2 * xq[qz] * pz[i];
            S = c.m * xq[qz] + 2.0 * c.nb * zx + c.ndb * zz * xq[qz];
        } else {
            const double zz = pz[0] * pz[0] + pz[1] * pz[1] + pz[2] * pz[2];
#pragma unroll
            for (int i = 0; i < 3; ++i) Fp[i] = c.a1 * pz[i];
            S = c.m * xq[qz] + c.nb * zz;
        }
        F0[qz] = w * (iJ[0] * Fp[0] + iJ[1] * Fp[1] + iJ[2] * Fp[2]);
        F1[qz] = w * (iJ[3] * Fp[0] + iJ[4] * Fp[1] + iJ[5] * Fp[2]);
        F2[qz] = w * (iJ[6] * Fp[0] + iJ[7] * Fp[1] + iJ[8] * Fp[2]);
        v[qz] = w * S;
    }

    // transposed gradient
    __syncthreads();
#pragma unroll
    for (int qz = 0; qz < QZ; ++qz) {
        sA[(qz * Q + ty) * Q + tx] = F0[qz];
        sB[(qz * Q + ty) * Q + tx] = F1[qz];
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < QZ; ++q) s += t.Gqz[q * QZ + qz] * F2[q];
        v[qz] += s;
    }
    __syncthreads();
#pragma unroll
    for (int qz = 0; qz < QZ; ++qz) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            s += tt.GTx[q] * sA[(qz * Q + ty) * Q + q];
            s += tt.GTy[q] * sB[(qz * Q + q) * Q + tx];
        }
        v[qz] += s;
    }

    // transposed interpolation
    __syncthreads();
#pragma unroll
    for (int qz = 0; qz < QZ; ++qz) sA[(qz * Q + ty) * Q + tx] = v[qz];
    __syncthreads();
    if (tx < D) {
#pragma unroll
        for (int qz = 0; qz < QZ; ++qz) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < Q; ++q) s += tt.BTx[q] * sA[(qz * Q + ty) * Q + q];
            sB[(qz * Q + ty) * D + tx] = s;
        }
    }
    __syncthreads();
    if (node && active) {
        double b[QZ];
#pragma unroll
        for (int qz = 0; qz < QZ; ++qz) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < Q; ++q) s += tt.BTy[q] * sB[(qz * Q + q) * D + tx];
            b[qz] = s;
        }
#pragma unroll
        for (int k = 0; k < DZ; ++k) {
            double s = 0.0;
#pragma unroll
            for (int qz = 0; qz < QZ; ++qz) s += t.Bz[qz * DZ + k] * b[qz];
            const int g = gi[k];
            if (g >= 0) atomicAdd(a.y + g, s);
            else if (!a.mask_out) atomicAdd(a.y + (~g), s);
                else if (a.diag_one) a.y[~g] = a.x[~g];   // every writer stores the same value
        }
    }
}

template <int Q> struct DefaultNE { static constexpr int value = Q == 2 ? 32 : Q == 3 ? 16 : Q == 4 ? 8 : Q == 5 ? 6 : 4; };

}  // namespace jots
