// Register-slab element kernel for hexes with a diagonal metric (axis-aligned bricks, any axis
// permutation) -- the fast path of the operator apply.
//
// The generic kernel (apply_generic.cuh) moves one shared-memory word per DFMA and is bound by the
// 128 B/clk shared-memory port.  Here the contractions are regrouped so that almost all of them
// run out of registers:
//   phase 1  one thread per (element, i, j) nodal pencil: gather the L-vector DOFs, evaluate the
//            nodal property polynomials in registers, contract along z with B and G, store the
//            resulting 3x3 planes for each of the Q quadrature planes;
//   phase 2  two mirrored threads per (element, quadrature plane qz): load the few planes, do every
//            in-plane interpolation / derivative, the quadrature-point physics and the transposed
//            in-plane contractions entirely in registers (about 800 FP64 instructions per thread for
//            the Q2 Jacobian apply), store two 3x3 result planes;
//   phase 3  one thread per (element, i, j): transposed z contraction, FP64 atomics to the L-vector.
// Shared-memory traffic drops from ~4400 to ~600 words per Q2 element.
#pragma once
#include "coef.cuh"

namespace jots {

template <int D, int Q>
struct SlabTables {
    double B[Q * D], G[Q * D];     // B[q*D+i] = l_i(x_q), G = derivative
    double BW[Q * D], GW[Q * D];   // rows scaled by the quadrature weight W[q]
    double W[Q];
};

template <int OP, int CF> struct SlabCfg {
    static constexpr bool X_GRAD = (OP != OP_RES);
    static constexpr bool Z_GRAD = (OP != OP_LIN);
    static constexpr int NF = CoefNF<CF, OP>::value;          // CF_CONST: 0, CF_K: 1 (2 for OP_JAC)
    static constexpr int NP = 1 + (X_GRAD ? 1 : 0) + (Z_GRAD ? 2 : 0) + NF;
    // plane ids
    static constexpr int P_XB = 0;
    static constexpr int P_XG = 1;                              // if X_GRAD
    static constexpr int P_ZB = 1 + (X_GRAD ? 1 : 0);           // if Z_GRAD
    static constexpr int P_ZG = P_ZB + 1;
    static constexpr int P_F0 = 1 + (X_GRAD ? 1 : 0) + (Z_GRAD ? 2 : 0);
};

template <int D> struct PlaneStride { static constexpr int value = (D * D + 1) / 2 * 2 + (((D * D + 1) / 2) % 2 == 0 ? 2 : 0); };
// D=2: 4 -> 4+2 = 6 ; D=3: 10 ; D=4: 16+2 = 18   (stride/2 odd => conflict-free 128-bit plane loads)

template <int D, int Q, int NP> struct ElemStride {
    static constexpr int PS = PlaneStride<D>::value;
    static constexpr int raw = (NP + 2) * Q * PS;
    static constexpr int value = raw + ((8 - raw % 16) + 16) % 16;   // == 8 (mod 16)
};

template <int D, int Q>
__device__ __forceinline__ void xpass(const double (&in)[D * D], const double* tab, double (&out)[D * Q]) {
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            double s = tab[q * D] * in[j * D];
#pragma unroll
            for (int i = 1; i < D; ++i) s += tab[q * D + i] * in[j * D + i];
            out[j * Q + q] = s;
        }
}
template <int D, int Q>
__device__ __forceinline__ void ypass(const double (&in)[D * Q], const double* tab, double (&out)[Q * Q]) {
#pragma unroll
    for (int qy = 0; qy < Q; ++qy)
#pragma unroll
        for (int qx = 0; qx < Q; ++qx) {
            double s = tab[qy * D] * in[qx];
#pragma unroll
            for (int j = 1; j < D; ++j) s += tab[qy * D + j] * in[j * Q + qx];
            out[qy * Q + qx] = s;
        }
}
// out[j][qx] = sum_qy tab[qy][j] in[qy][qx]
template <int D, int Q>
__device__ __forceinline__ void ypassT(const double (&in)[Q * Q], const double* tab, double (&out)[D * Q]) {
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int qx = 0; qx < Q; ++qx) {
            double s = tab[j] * in[qx];
#pragma unroll
            for (int qy = 1; qy < Q; ++qy) s += tab[qy * D + j] * in[qy * Q + qx];
            out[j * Q + qx] = s;
        }
}
// acc[j][i] += scale * sum_qx tab[qx][i] in[j][qx]
template <int D, int Q>
__device__ __forceinline__ void xpassT_acc(const double (&in)[D * Q], const double* tab, double scale, double (&acc)[D * D]) {
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double s = tab[i] * in[j * Q];
#pragma unroll
            for (int qx = 1; qx < Q; ++qx) s += tab[qx * D + i] * in[j * Q + qx];
            acc[j * D + i] += scale * s;
        }
}

template <int D, int PS>
__device__ __forceinline__ void load_plane(const double* p, double (&out)[D * D]) {
    const double2* p2 = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int m = 0; m < (D * D) / 2; ++m) { const double2 v = p2[m]; out[2 * m] = v.x; out[2 * m + 1] = v.y; }
    if ((D * D) % 2) out[D * D - 1] = p[D * D - 1];
}
template <int D, int PS>
__device__ __forceinline__ void store_plane(double* p, const double (&in)[D * D]) {
    double2* p2 = reinterpret_cast<double2*>(p);
#pragma unroll
    for (int m = 0; m < (D * D) / 2; ++m) p2[m] = make_double2(in[2 * m], in[2 * m + 1]);
    if ((D * D) % 2) p[D * D - 1] = in[D * D - 1];
}

// =================================================================================================
// v2: two threads per quadrature plane.  Thread h = 0 owns the quadrature columns qx < Q/2, thread
// h = 1 the mirrored columns Q-1-qx.  Gauss / Gauss-Lobatto points are symmetric about 1/2, so
// B[Q-1-q][D-1-i] = B[q][i] and G[Q-1-q][D-1-i] = -G[q][i]: the mirrored thread runs the *same*
// instruction stream with the same compile-time tables on the plane read with i reversed (the sign
// of G cancels because every x-derivative is contracted with G again).  Registers per thread halve,
// warps per SM double, no contraction is duplicated (for odd Q the middle column is shared, each
// thread taking half its weight).
template <int D, int Q>
struct Slab2Tables {
    double B[Q * D], G[Q * D];      // full tables (y / z passes)
    double BW[Q * D], GW[Q * D];    // W[q] * row (transposed y pass)
    double BWh[Q * D], GWh[Q * D];  // W[q] * row, middle column halved when Q is odd (transposed x pass)
    double W[Q];
};

// 64-bit shared-memory accesses: a half-warp must touch 16 distinct 8-byte slots.  Phase 1/3 lanes
// run over (element, ij) items -> element stride == D*D (mod 16); phase 2 lanes over (element, qz,
// h) -> plane stride chosen so that {qz*PS + el*ES} stay distinct (D=3: 10, D=2: 5).
template <int D> struct Plane2Stride { static constexpr int value = D == 3 ? 10 : D == 2 ? 5 : 17; };
template <int D, int Q, int NP> struct Elem2Stride {
    static constexpr int PS = Plane2Stride<D>::value;
    static constexpr int raw = (NP + 4) * Q * PS;
    static constexpr int value = raw + (((D * D) % 16 - raw % 16) + 16) % 16;   // == D*D (mod 16)
};

// Mirror symmetry of the 1-D tables (Gauss / Gauss-Lobatto points are symmetric about 1/2): row Q-1-q is row q read
// backwards, with a sign flip for derivatives.  SymTab serves the upper rows from the lower ones at compile time, so
// that a table has Q*D/2 distinct values and the whole working set (B, G, W B, W G: 24 doubles at Q2) stays in
// uniform registers instead of being re-fetched from the constant bank.  The host symmetrises the tables
// (fill_slab2_tables), so the aliasing changes no bit.
template <int D, int Q, int SIGN>
struct SymTab {
    const double* p;
    __device__ __forceinline__ double operator[](int idx) const {
        const int q = idx / D, i = idx - q * D;
        return (2 * q < Q) ? p[idx] : SIGN * p[(Q - 1 - q) * D + (D - 1 - i)];
    }
};

template <int D, int QC, class Tab>
__device__ __forceinline__ void xpassC(const double (&in)[D * D], const Tab tab, double (&out)[D * QC]) {
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int c = 0; c < QC; ++c) {
            double s = tab[c * D] * in[j * D];
#pragma unroll
            for (int i = 1; i < D; ++i) s += tab[c * D + i] * in[j * D + i];
            out[j * QC + c] = s;
        }
}
template <int D, int Q, int QC, class Tab>
__device__ __forceinline__ void ypassC(const double (&in)[D * QC], const Tab tab, double (&out)[Q * QC]) {
#pragma unroll
    for (int qy = 0; qy < Q; ++qy)
#pragma unroll
        for (int c = 0; c < QC; ++c) {
            double s = tab[qy * D] * in[c];
#pragma unroll
            for (int j = 1; j < D; ++j) s += tab[qy * D + j] * in[j * QC + c];
            out[qy * QC + c] = s;
        }
}
template <int D, int Q, int QC, class Tab>
__device__ __forceinline__ void ypassTC(const double (&in)[Q * QC], const Tab tab, double (&out)[D * QC]) {
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int c = 0; c < QC; ++c) {
            double s = tab[j] * in[c];
#pragma unroll
            for (int qy = 1; qy < Q; ++qy) s += tab[qy * D + j] * in[qy * QC + c];
            out[j * QC + c] = s;
        }
}
template <int D, int QC, class Tab>
__device__ __forceinline__ void xpassTC_acc(const double (&in)[D * QC], const Tab tab, double scale, double (&acc)[D * D]) {
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double s = tab[i] * in[j * QC];
#pragma unroll
            for (int c = 1; c < QC; ++c) s += tab[c * D + i] * in[j * QC + c];
            acc[j * D + i] += scale * s;
        }
}
// acc[j][i] += sum_c tab[c][i] in[j][c]  (input already scaled: one DFMA per term, straight into the accumulator)
template <int D, int QC, class Tab>
__device__ __forceinline__ void xpassTC_add(const double (&in)[D * QC], const Tab tab, double (&acc)[D * D]) {
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int c = 0; c < QC; ++c) acc[j * D + i] += tab[c * D + i] * in[j * QC + c];
}
// reverse the i index of a D x D plane when flip is set (branch-free selects)
template <int D>
__device__ __forceinline__ void flip_plane(double (&p)[D * D], bool flip) {
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int i = 0; i < D / 2; ++i) {
            const double lo = p[j * D + i], hi = p[j * D + D - 1 - i];
            p[j * D + i] = flip ? hi : lo;
            p[j * D + D - 1 - i] = flip ? lo : hi;
        }
}


// =================================================================================================
// Diagonal of an OP_LIN / OP_JAC form with the same three-phase structure: the diagonal of a
// tensor-product form is a transposed interpolation with squared tables,
//   d_i = sum_q w_q [ a1 sum_a g_aa (d_a phi_i)^2 + a2 phi_i sum_a g_aa d_a z d_a phi_i + m phi_i^2 ],
// (d_x phi_i)^2 = G^2[qx][i1] B^2[qy][i2] B^2[qz][i3] etc.  Replaces the generic loop kernel (41 ms at
// 17M DOFs) on axis-aligned hexes.
template <int D, int Q>
struct SlabDiagTables {
    double B[Q * D], G[Q * D];
    double B2W[Q * D], G2W[Q * D], BGW[Q * D];       // W[q] * B^2, W[q] * G^2, W[q] * B * G
    double B2Wh[Q * D], G2Wh[Q * D], BGWh[Q * D];    // middle column halved when Q is odd
    double B2[Q * D], G2[Q * D], BG[Q * D];          // unweighted (z direction; W[qz] is in wz)
    double mB[D];                                    // sum_q W[q] B[q][i]^2
    double W[Q];
};

template <int D, int Q, int OP, int CF, int E, int T>
__global__ void __launch_bounds__(T) form_diag_slab2_kernel(const FormArgs a, const SlabDiagTables<D, Q> t) {
    constexpr int NF = CoefNF<CF, OP>::value, NFA = NF > 0 ? NF : 1;
    constexpr bool JAC = (OP == OP_JAC);
    constexpr int NPI = NF + (JAC ? 2 : 0);          // input planes: fields, zB, zG
    constexpr int P_ZB = NF, P_ZG = NF + 1;
    constexpr int QC = (Q + 1) / 2;
    constexpr int PS = Plane2Stride<D>::value;
    constexpr int ES = Elem2Stride<D, Q, NPI + 2>::value;   // NPI + 6 planes
    constexpr int ND = D * D * D, DD = D * D;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int e0 = blockIdx.x * E;

    if (NPI > 0) {
        for (int item = tid; item < E * DD; item += T) {
            const int el = item / DD, ij = item - el * DD;
            const int e = e0 + el;
            if (e >= a.nelem) continue;
            double zr[D], fr[NFA][D];
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const int g = a.gather[(size_t)e * ND + ij + DD * k];
                zr[k] = a.z[g < 0 ? ~g : g];
                double f[NFA];
                nodal_fields<CF, OP>(a.mat, zr[k], f);
#pragma unroll
                for (int n = 0; n < NF; ++n) fr[n][k] = f[n];
            }
            double* base = smem + el * ES + ij;
#pragma unroll
            for (int qz = 0; qz < Q; ++qz) {
                if (JAC) {
                    double zs = 0.0, zg = 0.0;
#pragma unroll
                    for (int k = 0; k < D; ++k) { zs += t.B[qz * D + k] * zr[k]; zg += t.G[qz * D + k] * zr[k]; }
                    base[(P_ZB * Q + qz) * PS] = zs; base[(P_ZG * Q + qz) * PS] = zg;
                }
#pragma unroll
                for (int n = 0; n < NF; ++n) {
                    double fs = 0.0;
#pragma unroll
                    for (int k = 0; k < D; ++k) fs += t.B[qz * D + k] * fr[n][k];
                    base[(n * Q + qz) * PS] = fs;
                }
            }
        }
    }
    __syncthreads();

    if (tid < E * Q * 2) {
        const int el = tid / (2 * Q), r = tid - el * 2 * Q, qz = r >> 1;
        const bool h = r & 1;
        int e = e0 + el;
        if (e >= a.nelem) e = a.nelem - 1;
        const double* gm = a.geom + (size_t)e * a.geom_stride;
        const double gxx = gm[0] * gm[0] + gm[1] * gm[1] + gm[2] * gm[2];
        const double gyy = gm[3] * gm[3] + gm[4] * gm[4] + gm[5] * gm[5];
        const double gzz = gm[6] * gm[6] + gm[7] * gm[7] + gm[8] * gm[8];
        const double wz = t.W[qz] * gm[9];
        const double* pl = smem + el * ES + qz * PS;
        auto getp = [&](int id, double (&out)[DD]) {
            const double* p = pl + id * Q * PS;
#pragma unroll
            for (int m = 0; m < DD; ++m) out[m] = p[m];
            flip_plane<D>(out, h);
        };
        double p9[DD], t6[D * QC], A[DD], Bz[DD], Cz[DD];
#pragma unroll
        for (int i = 0; i < DD; ++i) { A[i] = 0.0; Bz[i] = 0.0; Cz[i] = 0.0; }
        double a1[Q * QC];
        if (CF == CF_CONST) {
#pragma unroll
            for (int q = 0; q < Q * QC; ++q) a1[q] = a.ck * a.a0;
        } else {
            getp(CF == CF_LINFIELDS ? 2 : 0, p9);
            xpassC<D, QC>(p9, t.B, t6);
            ypassC<D, Q, QC>(t6, t.B, a1);
            const double sc = CF == CF_LINFIELDS ? a.ck : a.ck * a.inv_rc;
#pragma unroll
            for (int q = 0; q < Q * QC; ++q) a1[q] *= sc;
        }
        if (CF == CF_LINFIELDS) {
            // mass diagonal with the interpolated rho_h * C_h
            double mq[Q * QC], cq[Q * QC];
            getp(0, p9);
            xpassC<D, QC>(p9, t.B, t6);
            ypassC<D, Q, QC>(t6, t.B, mq);
            getp(1, p9);
            xpassC<D, QC>(p9, t.B, t6);
            ypassC<D, Q, QC>(t6, t.B, cq);
#pragma unroll
            for (int q = 0; q < Q * QC; ++q) mq[q] *= cq[q];
            ypassTC<D, Q, QC>(mq, t.B2W, t6);
            xpassTC_acc<D, QC>(t6, t.B2Wh, a.cm * wz, A);
        }
        ypassTC<D, Q, QC>(a1, t.B2W, t6);
        xpassTC_acc<D, QC>(t6, t.G2Wh, wz * gxx, A);
        xpassTC_acc<D, QC>(t6, t.B2Wh, wz * gzz, Bz);
        ypassTC<D, Q, QC>(a1, t.G2W, t6);
        xpassTC_acc<D, QC>(t6, t.B2Wh, wz * gyy, A);
        if (JAC) {
            double a2[Q * QC], zB6[D * QC], zG6[D * QC], F[Q * QC];
            getp(1, p9);
            xpassC<D, QC>(p9, t.B, t6);
            ypassC<D, Q, QC>(t6, t.B, a2);
            const double sc = a.ck * a.inv_rc;
#pragma unroll
            for (int q = 0; q < Q * QC; ++q) a2[q] *= sc;
            getp(P_ZB, p9);
            xpassC<D, QC>(p9, t.B, zB6);
            xpassC<D, QC>(p9, t.G, zG6);
            ypassC<D, Q, QC>(zG6, t.B, F);
#pragma unroll
            for (int q = 0; q < Q * QC; ++q) F[q] *= a2[q];
            ypassTC<D, Q, QC>(F, t.B2W, t6);
            xpassTC_acc<D, QC>(t6, t.BGWh, wz * gxx, A);
            ypassC<D, Q, QC>(zB6, t.G, F);
#pragma unroll
            for (int q = 0; q < Q * QC; ++q) F[q] *= a2[q];
            ypassTC<D, Q, QC>(F, t.BGW, t6);
            xpassTC_acc<D, QC>(t6, t.B2Wh, wz * gyy, A);
            getp(P_ZG, p9);
            xpassC<D, QC>(p9, t.B, t6);
            ypassC<D, Q, QC>(t6, t.B, F);
#pragma unroll
            for (int q = 0; q < Q * QC; ++q) F[q] *= a2[q];
            ypassTC<D, Q, QC>(F, t.B2W, t6);
            xpassTC_acc<D, QC>(t6, t.B2Wh, wz * gzz, Cz);
        }
        flip_plane<D>(A, h); flip_plane<D>(Bz, h); flip_plane<D>(Cz, h);
        double* outp = smem + el * ES + ((NPI + (h ? 3 : 0)) * Q + qz) * PS;
#pragma unroll
        for (int m = 0; m < DD; ++m) { outp[m] = A[m]; outp[Q * PS + m] = Bz[m]; outp[2 * Q * PS + m] = Cz[m]; }
    }
    __syncthreads();

    for (int item = tid; item < E * DD; item += T) {
        const int el = item / DD, ij = item - el * DD;
        const int e = e0 + el;
        if (e >= a.nelem) continue;
        const double* base = smem + el * ES + NPI * Q * PS + ij;
        const double* gm = a.geom + (size_t)e * a.geom_stride;
        const double mass = (CF == CF_LINFIELDS ? 0.0 : a.cm * a.m0) * gm[9] * t.mB[ij % D] * t.mB[ij / D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            double s = mass * t.mB[k];
#pragma unroll
            for (int qz = 0; qz < Q; ++qz)
                s += t.B2[qz * D + k] * (base[qz * PS] + base[(3 * Q + qz) * PS]) +
                     t.G2[qz * D + k] * (base[(Q + qz) * PS] + base[(4 * Q + qz) * PS]) +
                     t.BG[qz * D + k] * (base[(2 * Q + qz) * PS] + base[(5 * Q + qz) * PS]);
            const int g = a.gather[(size_t)e * ND + ij + DD * k];
            if (g >= 0) atomicAdd(a.y + g, s);
            else if (!a.mask_out) atomicAdd(a.y + (~g), s);
                else if (a.diag_one) a.y[~g] = a.x[~g];   // every writer stores the same value
        }
    }
}

// =================================================================================================
// v3: persistent, software-pipelined version of v2.  Each block loops over tiles; while it computes
// tile i (phase 2, FP64 bound) the L-vector values of tile i+1 and the gather rows of tile i+2 are in
// flight as cp.async copies into shared-memory staging buffers, so no warp waits on a global load
// (the long-scoreboard stalls that took 35 % of v2's issue slots).  The two result planes of each
// half-plane thread alias the input planes of the same (element, qz) -- both threads of the pair are
// adjacent lanes of one warp, a __syncwarp() orders their last read before the first write -- which
// keeps four 128-thread blocks (16 warps) resident per SM.
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// v3 shared-memory layout (found by exhaustive search for conflict-free 64-bit accesses of a half-warp):
//   plane (id, qz) of an element starts at id*PSET + qz*PS; elements are ES apart.
//   MIRROR: thread h=1 reads the planes with i reversed through its address (no register selects).
template <int D, int Q> struct Slab3Layout {   // D = 2: no conflict-free mirrored layout exists, keep selects
    static constexpr int PS = 5, PAD = 0, ESM = 4; static constexpr bool MIRROR = false;
};
template <> struct Slab3Layout<3, 4> { static constexpr int PS = 12, PAD = 2, ESM = 9; static constexpr bool MIRROR = true; };
// Q3: mirrored addressing spills at 255 registers; (PS, PAD, ESM) = (17, 3, 13) makes the phase-2 plane loads and result stores of a
// half-warp (10 lanes per element) conflict free -- (17, 0, 0) of round 1 took 1.9 wavefronts for each (profiles/r2p_phases_slab3_q3_*)
template <> struct Slab3Layout<4, 5> { static constexpr int PS = 17, PAD = 3, ESM = 13; static constexpr bool MIRROR = false; };
template <int D, int Q, int NP> struct Elem3Stride {
    static constexpr int PS = Slab3Layout<D, Q>::PS;
    static constexpr int PSET = Q * PS + Slab3Layout<D, Q>::PAD;
    static constexpr int NPA = NP > 4 ? NP : 4;                       // room for the 4 aliased result planes
    static constexpr int raw = NPA * PSET;
    static constexpr int value = raw + ((Slab3Layout<D, Q>::ESM - raw % 16) + 16) % 16;
};
// Staging buffers for the L-vector values: two (values of tile i+1 land while tile i is in phases 1-3), or one when the
// tile is so large that a second buffer would cost a resident block (Q2, E = 16: 52 KB instead of 59 KB per block keeps
// four blocks per SM); the single buffer is refilled right after phase 1 and has phases 2-3 to land.
template <int Q, int E> struct Slab3Stages { static constexpr int value = (Q == 4 && E >= 16) ? 1 : 2; };
template <int D, int Q, int NP, int E> struct Slab3Smem {
    static constexpr int ND = D * D * D;
    static constexpr size_t planes = (size_t)E * Elem3Stride<D, Q, NP>::value * sizeof(double);
    static constexpr size_t stage = (size_t)Slab3Stages<Q, E>::value * 2 * E * ND * sizeof(double);
    static constexpr size_t idx = (size_t)3 * E * ND * sizeof(int);
    static constexpr size_t total = planes + stage + idx;
};

// Phase 2 of the pipelined slab kernels: one thread owns half a quadrature plane (element, qz, h) and keeps every
// in-plane contraction and the quadrature-point physics in registers.  pl points at plane (id 0, qz) of the
// element; plane id sits at pl + id*PSET.  The two result planes alias input planes 0..3 of the same (element,
// qz): only this thread pair (adjacent lanes) ever touches them, __syncwarp orders last read before first write.
// RS is the row stride of a plane (D in the per-element layout; the node-row length of a patch in apply_patch.cuh).
// ALIAS = false: the results go to outA / outB (D*D contiguous doubles each) instead of the aliased input planes.
template <int D, int Q, int OP, int CF, int PSET, bool MIRROR, int RS = D, bool ALIAS = true>
__device__ __forceinline__ void slab_phase2(const FormArgs& a, const Slab2Tables<D, Q>& t, double* pl, const bool h,
                                            const double wz, const double gxx, const double gyy, const double gzz,
                                            const unsigned p2mask, double* outA = nullptr, double* outB = nullptr) {
    typedef SlabCfg<OP, CF> C;
    constexpr int QC = (Q + 1) / 2;
    constexpr int DD = D * D;
    // mirrored rows come from the lower half of each table; for even Q the half-weight tables are the full-weight ones
    const SymTab<D, Q, 1> tB{t.B}, tBW{t.BW}, tBWh{Q % 2 == 0 ? t.BW : t.BWh};
    const SymTab<D, Q, -1> tG{t.G}, tGW{t.GW}, tGWh{Q % 2 == 0 ? t.GW : t.GWh};
    int mo[D];   // mirrored column offsets of the h = 1 thread
#pragma unroll
    for (int i = 0; i < D; ++i) mo[i] = (MIRROR && h) ? D - 1 - i : i;
    auto getp = [&](int id, double (&out)[DD]) {
        const double* p = pl + id * PSET;
        if (MIRROR) {
#pragma unroll
            for (int jj = 0; jj < D; ++jj)
#pragma unroll
                for (int i = 0; i < D; ++i) out[jj * D + i] = p[jj * RS + mo[i]];
        } else {
#pragma unroll
            for (int m = 0; m < DD; ++m) out[m] = p[(m / D) * RS + (m % D)];
            flip_plane<D>(out, h);
        }
    };
    double p9[DD], t6[D * QC], A[DD], Bp[DD];
#pragma unroll
    for (int i = 0; i < DD; ++i) { A[i] = 0.0; Bp[i] = 0.0; }
    if constexpr (CF == CF_FULL || CF == CF_FULLC) {
        // rho(T) and/or C(T): alpha, alpha', -beta, -beta' of nl_conduction_operator.cpp:92-163 from the
        // interpolated nodal fields (k, k', rho, rho', rho'', C, C', C''; CF_FULLC: uniform rho, five
        // fields), with 1/rho and 1/C formed once per quadrature point
        constexpr int NQH = Q * QC;
        constexpr bool RHO = (CF == CF_FULL);
        constexpr int F_K = 0, F_DK = 1, F_R = 2, F_DR = 3, F_D2R = 4, F_C = RHO ? 5 : 2, F_DC = RHO ? 6 : 3, F_D2C = RHO ? 7 : 4;
        auto field = [&](int n, double (&out)[NQH]) {
            getp(C::P_F0 + n, p9);
            xpassC<D, QC>(p9, tB, t6);
            ypassC<D, Q, QC>(t6, tB, out);
        };
        double a1[NQH], a2[NQH], nb[NQH], ndb[NQH];
        {
            double ir[NQH], ic[NQH], sr[NQH], sc[NQH], tmpq[NQH];
            const double ir0 = 1.0 / a.mat.rho.c[0];
            field(F_C, ic);
            if (RHO) field(F_R, ir);
#pragma unroll
            for (int q = 0; q < NQH; ++q) { ir[q] = RHO ? 1.0 / ir[q] : ir0; ic[q] = 1.0 / ic[q]; }
            field(F_DC, sc);
            if (RHO) field(F_DR, sr);
#pragma unroll
            for (int q = 0; q < NQH; ++q) { sr[q] = RHO ? sr[q] * ir[q] : 0.0; sc[q] *= ic[q]; }
            if (OP == OP_JAC) {
                // inner = irc (-C''/C + 2 sC sR + 2 sC^2 - rho''/rho + 2 sR^2)
                field(F_D2C, tmpq);
#pragma unroll
                for (int q = 0; q < NQH; ++q) ndb[q] = 2.0 * (sc[q] * sr[q] + sc[q] * sc[q] + sr[q] * sr[q]) - tmpq[q] * ic[q];
                if (RHO) {
                    field(F_D2R, tmpq);
#pragma unroll
                    for (int q = 0; q < NQH; ++q) ndb[q] -= tmpq[q] * ir[q];
                }
            }
            field(F_K, a1);
#pragma unroll
            for (int q = 0; q < NQH; ++q) {
                const double irc = ir[q] * ic[q], sum = sr[q] + sc[q], kq = a1[q];
                a1[q] = a.ck * kq * irc;
                nb[q] = a.cb * kq * irc * sum;                         // cb * (-beta)
                if (OP == OP_JAC) {
                    ndb[q] = kq * irc * ndb[q];                        // k * inner
                    a2[q] = kq * sum;                                  // finished below with k'
                    sr[q] = irc; sc[q] = sum;
                }
            }
            if (OP == OP_JAC) {
                field(F_DK, tmpq);
#pragma unroll
                for (int q = 0; q < NQH; ++q) {
                    const double irc = sr[q], sum = sc[q];
                    a2[q] = a.ck * irc * (tmpq[q] - a2[q]);              // ck * alpha'
                    ndb[q] = a.cb * (tmpq[q] * irc * sum - ndb[q]);      // cb * (-beta') = -cb (k' t + k inner), t = -irc sum
                }
            }
        }
        double xB6[D * QC], xG6[D * QC], zB6[D * QC], zG6[D * QC], S[NQH];
        getp(C::P_XB, p9);
        xpassC<D, QC>(p9, tB, xB6);
        if (C::X_GRAD) xpassC<D, QC>(p9, tG, xG6);
        ypassC<D, Q, QC>(xB6, tB, S);                                   // x at the quadrature points
        if (OP == OP_JAC) {
#pragma unroll
            for (int q = 0; q < NQH; ++q) { a2[q] *= S[q]; ndb[q] *= S[q]; }   // alpha' x, (-beta') x
        }
#pragma unroll
        for (int q = 0; q < NQH; ++q) S[q] *= a.cm;
        getp(C::P_ZB, p9);
        xpassC<D, QC>(p9, tB, zB6);
        xpassC<D, QC>(p9, tG, zG6);
        // one direction at a time: F_d = a1 xd + (alpha' x) zd (Jacobian) | a1 zd (residual);
        // S += g_dd (2 nb zd xd + (-beta') x zd^2) | g_dd nb zd^2
        auto direction = [&](int dir, double gdd, double (&acc)[DD]) {
            double F[NQH], zd[NQH];
            if (dir == 2) {
                getp(C::P_ZG, p9);
                xpassC<D, QC>(p9, tB, t6);
                ypassC<D, Q, QC>(t6, tB, zd);
                if (OP == OP_JAC) {
                    getp(C::P_XG, p9);
                    xpassC<D, QC>(p9, tB, t6);
                    ypassC<D, Q, QC>(t6, tB, F);
                }
            } else {
                if (dir == 0) ypassC<D, Q, QC>(zG6, tB, zd);
                else ypassC<D, Q, QC>(zB6, tG, zd);
                if (OP == OP_JAC) {
                    if (dir == 0) ypassC<D, Q, QC>(xG6, tB, F);
                    else ypassC<D, Q, QC>(xB6, tG, F);
                }
            }
#pragma unroll
            for (int q = 0; q < NQH; ++q) {
                if (OP == OP_JAC) {
                    S[q] += gdd * zd[q] * (2.0 * nb[q] * F[q] + ndb[q] * zd[q]);
                    F[q] = a1[q] * F[q] + a2[q] * zd[q];
                } else {
                    S[q] += gdd * nb[q] * zd[q] * zd[q];
                    F[q] = a1[q] * zd[q];
                }
            }
            if (dir == 1) ypassTC<D, Q, QC>(F, tGW, t6);
            else ypassTC<D, Q, QC>(F, tBW, t6);
            if (dir == 0) xpassTC_acc<D, QC>(t6, tGWh, wz * gdd, acc);
            else xpassTC_acc<D, QC>(t6, tBWh, wz * gdd, acc);
        };
        direction(0, gxx, A);
        direction(1, gyy, A);
        direction(2, gzz, Bp);
        ypassTC<D, Q, QC>(S, tBW, t6);
        xpassTC_acc<D, QC>(t6, tBWh, wz, A);
    } else {
        double a1[Q * QC], a2x[Q * QC];
        double a1c = 0.0;
        if (CF == CF_CONST) a1c = a.ck * a.a0;
        else {
            getp(C::P_F0 + (CF == CF_LINFIELDS ? 2 : 0), p9);
            xpassC<D, QC>(p9, tB, t6);
            ypassC<D, Q, QC>(t6, tB, a1);
            if (CF == CF_LINFIELDS) {
                // mass coefficient rho_h * C_h at the quadrature points (ProductCoefficient of two
                // interpolated fields, linear_conduction_operator.cpp:20,59) -> a2x
                double cq[Q * QC];
                getp(C::P_F0, p9);
                xpassC<D, QC>(p9, tB, t6);
                ypassC<D, Q, QC>(t6, tB, a2x);
                getp(C::P_F0 + 1, p9);
                xpassC<D, QC>(p9, tB, t6);
                ypassC<D, Q, QC>(t6, tB, cq);
    #pragma unroll
                for (int q = 0; q < Q * QC; ++q) a2x[q] *= cq[q];
            }
            if (OP == OP_JAC && CF != CF_KL) {
                getp(C::P_F0 + 1, p9);
                xpassC<D, QC>(p9, tB, t6);
                ypassC<D, Q, QC>(t6, tB, a2x);
            }
        }
        double xB6[D * QC], xG6[D * QC], vy6[D * QC];
        getp(C::P_XB, p9);
        xpassC<D, QC>(p9, tB, xB6);
        if (C::X_GRAD) xpassC<D, QC>(p9, tG, xG6);
        {
            double xv[Q * QC];
            ypassC<D, Q, QC>(xB6, tB, xv);
            if (OP == OP_JAC && CF == CF_KL) {
                // k linear in T: the interpolant of the nodal k' is the uniform leading coefficient
                const double a2c = a.ck * a.inv_rc * a.mat.k.c[0];
    #pragma unroll
                for (int q = 0; q < Q * QC; ++q) a2x[q] = a2c * xv[q];
            } else if (OP == OP_JAC && CF != CF_CONST) {
    #pragma unroll
                for (int q = 0; q < Q * QC; ++q) a2x[q] *= xv[q];
            }
            if (CF == CF_LINFIELDS) {
    #pragma unroll
                for (int q = 0; q < Q * QC; ++q) xv[q] *= a2x[q];
            }
            ypassTC<D, Q, QC>(xv, tBW, vy6);
        }
        const double c_m = a.cm * (CF == CF_CONST || CF == CF_K || CF == CF_KL ? a.m0 : 1.0) * wz;
        double zB6[D * QC], zG6[D * QC];
        if (C::Z_GRAD) {
            getp(C::P_ZB, p9);
            xpassC<D, QC>(p9, tB, zB6);
            xpassC<D, QC>(p9, tG, zG6);
        }
        {   // x direction (both G's carry the mirror sign: it cancels)
            double F[Q * QC];
            if (OP == OP_RES) ypassC<D, Q, QC>(zG6, tB, F);
            else ypassC<D, Q, QC>(xG6, tB, F);
            if (CF == CF_CONST) {
    #pragma unroll
                for (int q = 0; q < Q * QC; ++q) F[q] *= a1c;
            } else {
    #pragma unroll
                for (int q = 0; q < Q * QC; ++q) F[q] *= a1[q];
                if (OP == OP_JAC) {
                    double zd[Q * QC];
                    ypassC<D, Q, QC>(zG6, tB, zd);
    #pragma unroll
                    for (int q = 0; q < Q * QC; ++q) F[q] += a2x[q] * zd[q];
                }
            }
            ypassTC<D, Q, QC>(F, tBW, t6);
            {
                const double cx = wz * gxx;
#pragma unroll
                for (int i = 0; i < D * QC; ++i) t6[i] *= cx;
            }
            xpassTC_add<D, QC>(t6, tGWh, A);
        }
        {   // y direction (+ mass)
            double F[Q * QC];
            if (OP == OP_RES) ypassC<D, Q, QC>(zB6, tG, F);
            else ypassC<D, Q, QC>(xB6, tG, F);
            if (CF == CF_CONST) {
    #pragma unroll
                for (int q = 0; q < Q * QC; ++q) F[q] *= a1c;
            } else {
    #pragma unroll
                for (int q = 0; q < Q * QC; ++q) F[q] *= a1[q];
                if (OP == OP_JAC) {
                    double zd[Q * QC];
                    ypassC<D, Q, QC>(zB6, tG, zd);
    #pragma unroll
                    for (int q = 0; q < Q * QC; ++q) F[q] += a2x[q] * zd[q];
                }
            }
            ypassTC<D, Q, QC>(F, tGW, t6);
            const double cy = wz * gyy;
    #pragma unroll
            for (int i = 0; i < D * QC; ++i) t6[i] = cy * t6[i] + c_m * vy6[i];
            xpassTC_add<D, QC>(t6, tBWh, A);
        }
        {   // z direction
            double F[Q * QC];
            getp(OP == OP_RES ? C::P_ZG : C::P_XG, p9);
            xpassC<D, QC>(p9, tB, t6);
            ypassC<D, Q, QC>(t6, tB, F);
            if (CF == CF_CONST) {
    #pragma unroll
                for (int q = 0; q < Q * QC; ++q) F[q] *= a1c;
            } else {
    #pragma unroll
                for (int q = 0; q < Q * QC; ++q) F[q] *= a1[q];
                if (OP == OP_JAC) {
                    double zd[Q * QC];
                    getp(C::P_ZG, p9);
                    xpassC<D, QC>(p9, tB, t6);
                    ypassC<D, Q, QC>(t6, tB, zd);
    #pragma unroll
                    for (int q = 0; q < Q * QC; ++q) F[q] += a2x[q] * zd[q];
                }
            }
            ypassTC<D, Q, QC>(F, tBW, t6);
            {
                const double cz = wz * gzz;
#pragma unroll
                for (int i = 0; i < D * QC; ++i) t6[i] *= cz;
            }
            xpassTC_add<D, QC>(t6, tBWh, Bp);
        }
    }
    flip_plane<D>(A, h);
    flip_plane<D>(Bp, h);
    // result planes alias input planes [0..3][qz] of this (element, qz): only this thread pair
    // (adjacent lanes) ever touches them
    if constexpr (ALIAS) {
        __syncwarp(p2mask);
        // A -> plane h, Bp -> plane 2 + h (layout chosen so that the pair's stores do not collide)
        double* outp = pl + (h ? 1 : 0) * PSET;
#pragma unroll
        for (int m = 0; m < DD; ++m) { outp[m] = A[m]; outp[2 * PSET + m] = Bp[m]; }
    } else {
#pragma unroll
        for (int m = 0; m < DD; ++m) { outA[m] = A[m]; outB[m] = Bp[m]; }
    }
}

template <int D, int Q, int OP, int CF, int E, int T>
__global__ void __launch_bounds__(T) form_apply_slab3_kernel(const FormArgs a, const Slab2Tables<D, Q> t) {
    if (a.skip && *a.skip != 0.0) return;   // enqueued ahead of a Krylov solve that has already stopped
    typedef SlabCfg<OP, CF> C;
    constexpr int NF = C::NF, NFA = NF > 0 ? NF : 1, NP = C::NP;
    constexpr int PS = Elem3Stride<D, Q, NP>::PS, PSET = Elem3Stride<D, Q, NP>::PSET;
    constexpr int ES = Elem3Stride<D, Q, NP>::value;
    constexpr bool MIRROR = Slab3Layout<D, Q>::MIRROR;
    constexpr int ND = D * D * D, DD = D * D;
    constexpr bool NEED_Z = C::Z_GRAD || NF > 0;
    extern __shared__ __align__(16) double smem[];
    double* planes = smem;
    constexpr int NSTG = Slab3Stages<Q, E>::value;
    double* stage = smem + (size_t)E * ES;                                  // [NSTG][2][E*ND]
    int* idxbuf = reinterpret_cast<int*>(stage + (size_t)NSTG * 2 * E * ND);   // [3][E*ND]
    const int tid = threadIdx.x;
    const int ntiles = (a.nelem + E - 1) / E;
    const int nmine = blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    auto tile_of = [&](int i) { return (int)blockIdx.x + i * (int)gridDim.x; };
    auto nvalid = [&](int tile) { const int r = a.nelem - tile * E; return r < E ? r : E; };
    // gather rows of tile i -> idxbuf[i % 3]
    auto issue_idx = [&](int i) {
        if (i < nmine) {
            const int tile = tile_of(i), n = nvalid(tile) * ND;
            const int* src = a.gather + (size_t)tile * E * ND;
            int* dst = idxbuf + (i % 3) * E * ND;
#pragma unroll 1
            for (int m = tid; m < n; m += T) cp_async4(dst + m, src + m);
        }
        cp_async_commit();
    };
    // L-vector values of tile i (indices already in shared memory) -> stage[i % 2]
    auto issue_vals = [&](int i) {
        if (i < nmine) {
            const int n = nvalid(tile_of(i)) * ND;
            const int* ix = idxbuf + (i % 3) * E * ND;
            double* sx = stage + (size_t)(i % NSTG) * 2 * E * ND;
            double* sz = sx + E * ND;
#pragma unroll 1
            for (int m = tid; m < n; m += T) {
                const int g = ix[m];
                const int idx = g < 0 ? ~g : g;
                cp_async8(sx + m, a.x + idx);
                if (NEED_Z) cp_async8(sz + m, a.z + idx);
            }
        }
        cp_async_commit();
    };

    issue_idx(0);
    issue_idx(1);
    cp_async_wait<1>();
    __syncthreads();
    issue_vals(0);

    // FormArgs::dot (the conjugate-gradient loop's (d, A d) = sum over elements of d_e . (A_e d_e)): phase 3 multiplies every
    // row it adds to y with the staged input of that row.  Needs the values of tile `it` intact during its phase 3, i.e. two
    // staging buffers (the launcher offers the fused inner product only for those instantiations).
    const bool with_dot = NSTG == 2 && a.dot != nullptr;
    double dsum = 0.0;

    // geometry of a uniform mesh is loop invariant
    double gxx = 0, gyy = 0, gzz = 0, det = 0;
    if (a.geom_stride == 0) {
        const double* gm = a.geom;
        gxx = gm[0] * gm[0] + gm[1] * gm[1] + gm[2] * gm[2];
        gyy = gm[3] * gm[3] + gm[4] * gm[4] + gm[5] * gm[5];
        gzz = gm[6] * gm[6] + gm[7] * gm[7] + gm[8] * gm[8];
        det = gm[9];
    }

    for (int it = 0; it < nmine; ++it) {
        const int tile = tile_of(it), e0 = tile * E, nv = nvalid(tile);
        issue_idx(it + 2);
        cp_async_wait<1>();          // values of tile it and gather rows of tile it+1 have landed
        __syncthreads();
        if (NSTG == 2) issue_vals(it + 1);
        const int* ix = idxbuf + (it % 3) * E * ND;
        const double* sx = stage + (size_t)(it % NSTG) * 2 * E * ND;
        const double* sz = sx + E * ND;

        // ---------------- phase 1: z contraction from the staged values ---------------------------
#pragma unroll 1
        for (int item = tid; item < nv * DD; item += T) {
            const int el = item / DD, ij = item - el * DD;
            double xr[D], zr[D], fr[NFA][D];
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const int m = el * ND + ij + DD * k;
                const int g = ix[m];
                xr[k] = (g < 0 && a.mask_in) ? 0.0 : sx[m];
                if (NEED_Z) {
                    zr[k] = sz[m];
                    if (NF > 0) {
                        double f[NFA];
                        nodal_fields<CF, OP>(a.mat, zr[k], f);
#pragma unroll
                        for (int n = 0; n < NF; ++n)
                            fr[n][k] = (CF == CF_K || CF == CF_KL) ? (a.ck * a.inv_rc) * f[n] : (CF == CF_LINFIELDS && n == 2) ? a.ck * f[n] : f[n];
                    }
                }
            }
            double* base = planes + el * ES + ij;
#pragma unroll
            for (int qz = 0; qz < Q; ++qz) {
                double s = 0.0, sg = 0.0, zs = 0.0, zg = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    s += t.B[qz * D + k] * xr[k];
                    if (C::X_GRAD) sg += t.G[qz * D + k] * xr[k];
                    if (C::Z_GRAD) { zs += t.B[qz * D + k] * zr[k]; zg += t.G[qz * D + k] * zr[k]; }
                }
                base[C::P_XB * PSET + qz * PS] = s;
                if (C::X_GRAD) base[C::P_XG * PSET + qz * PS] = sg;
                if (C::Z_GRAD) { base[C::P_ZB * PSET + qz * PS] = zs; base[C::P_ZG * PSET + qz * PS] = zg; }
#pragma unroll
                for (int n = 0; n < NF; ++n) {
                    double fs = 0.0;
#pragma unroll
                    for (int k = 0; k < D; ++k) fs += t.B[qz * D + k] * fr[n][k];
                    base[(C::P_F0 + n) * PSET + qz * PS] = fs;
                }
            }
        }
        __syncthreads();
        if (NSTG == 1) issue_vals(it + 1);   // the staging buffer is free again; the copies land during phases 2-3

        // ---------------- phase 2: half a quadrature plane per thread ----------------------------
        const unsigned p2mask = __ballot_sync(0xffffffffu, tid < E * Q * 2);
        if (tid < E * Q * 2) {
            const int el = tid / (2 * Q), r = tid - el * 2 * Q, qz = r >> 1;
            const bool h = r & 1;
            if (a.geom_stride != 0) {
                int e = e0 + el;
                if (e >= a.nelem) e = a.nelem - 1;
                const double* gm = a.geom + (size_t)e * a.geom_stride;
                gxx = gm[0] * gm[0] + gm[1] * gm[1] + gm[2] * gm[2];
                gyy = gm[3] * gm[3] + gm[4] * gm[4] + gm[5] * gm[5];
                gzz = gm[6] * gm[6] + gm[7] * gm[7] + gm[8] * gm[8];
                det = gm[9];
            }
            const double wz = t.W[qz] * det;
            double* pl = planes + el * ES + qz * PS;
            slab_phase2<D, Q, OP, CF, PSET, MIRROR>(a, t, pl, h, wz, gxx, gyy, gzz, p2mask);
        }
        __syncthreads();

        // Side stream of the conjugate-gradient loop (FormArgs::side_*, see apply_patch2.cuh): this tile's slice of
        // x_sol += alpha d_old and of the zeroing of the next apply's output, at most two pairs of entries per thread.  The
        // loads are issued before phase 3 and consumed after it.
        double2 side_xv[2], side_dv[2];
        double side_al = 0.0;
        const long long side_k0 = (long long)tile * a.side_chunk;
        if (a.side_chunk != 0) {
            side_al = __ldg(a.side_alpha);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int o = 2 * (tid + r * T);
                const long long k = side_k0 + o;
                side_xv[r] = make_double2(0.0, 0.0); side_dv[r] = side_xv[r];
                if (o < a.side_chunk) {
                    if (k + 1 < a.nvec) {
                        side_xv[r] = __ldcs(reinterpret_cast<const double2*>(a.side_x + k));
                        side_dv[r] = __ldcs(reinterpret_cast<const double2*>(a.side_d + k));
                    } else if (k < a.nvec) { side_xv[r].x = __ldcs(a.side_x + k); side_dv[r].x = __ldcs(a.side_d + k); }
                }
            }
        }

        // ---------------- phase 3: transposed z contraction + scatter ----------------------------
#pragma unroll 1
        for (int item = tid; item < nv * DD; item += T) {
            const int el = item / DD, ij = item - el * DD;
            const double* base = planes + el * ES + ij;
            double Aq[Q], Bq[Q];
#pragma unroll
            for (int qz = 0; qz < Q; ++qz) {
                Aq[qz] = base[qz * PS] + base[PSET + qz * PS];
                Bq[qz] = base[2 * PSET + qz * PS] + base[3 * PSET + qz * PS];
            }
#pragma unroll
            for (int k = 0; k < D; ++k) {
                double s = 0.0;
#pragma unroll
                for (int qz = 0; qz < Q; ++qz) s += t.B[qz * D + k] * Aq[qz] + t.G[qz * D + k] * Bq[qz];
                const int m = el * ND + ij + DD * k;
                const int g = ix[m];
                if (g >= 0) atomicAdd(a.y + g, s);
                else if (!a.mask_out) atomicAdd(a.y + (~g), s);
                else if (a.diag_one) a.y[~g] = a.x[~g];   // every writer stores the same value
                // rows that are not written have a zero input (the caller fuses only forms with mask_in == mask_out)
                if (with_dot) dsum = fma((g < 0 && a.mask_in) ? 0.0 : sx[m], s, dsum);
            }
        }
        if (a.side_chunk != 0) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int o = 2 * (tid + r * T);
                const long long k = side_k0 + o;
                if (o < a.side_chunk) {
                    if (k + 1 < a.nvec) {
                        side_xv[r].x = fma(side_al, side_dv[r].x, side_xv[r].x);
                        side_xv[r].y = fma(side_al, side_dv[r].y, side_xv[r].y);
                        __stcs(reinterpret_cast<double2*>(a.side_x + k), side_xv[r]);
                        __stcs(reinterpret_cast<double2*>(a.side_zero + k), make_double2(0.0, 0.0));
                    } else if (k < a.nvec) {
                        __stcs(a.side_x + k, fma(side_al, side_dv[r].x, side_xv[r].x));
                        __stcs(a.side_zero + k, 0.0);
                    }
                }
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
    if (with_dot) {
        __shared__ double s_dsum[T / 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
        if ((tid & 31) == 0) s_dsum[tid >> 5] = dsum;
        __syncthreads();
        if (tid == 0) {
            double v = s_dsum[0];
#pragma unroll
            for (int w = 1; w < T / 32; ++w) v += s_dsum[w];
            atomicAdd(a.dot, v);
        }
    }
}

}  // namespace jots
