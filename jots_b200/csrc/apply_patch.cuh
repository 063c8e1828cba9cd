// Unique-pencil slab kernel: the three-phase register-slab scheme of apply_slab.cuh on PATCHES of PX x PY x 1
// lattice-adjacent hexes (patches.cpp) instead of runs of independent elements.
//
//   * a nodal pencil (the D nodes above one (I, J) position of the patch) that two or four elements share is gathered,
//     evaluated (property polynomials), contracted along z and scattered ONCE: a 4 x 4 patch of Q2 hexes has 9 x 9 = 81
//     pencils instead of 16 x 9 = 144, so phases 1 and 3 are one pass of three warps (the gather-table kernel runs a
//     second pass for 16 left-over pencils that every other warp waits for at the barrier: 20 % of its warp time),
//     with 44 % fewer cp.async gathers, polynomial evaluations and FP64 atomics;
//   * the quadrature planes are stored per patch, [plane id][qz][J][I]: phase 1 writes each value once, the two mirrored
//     threads of an (element, qz) pair read their 3 x 3 window at a compile-time offset;
//   * the last warp of the block never runs phases 1 / 3: it is the block's copy engine (patch descriptors two tiles ahead,
//     L-vector values one tile ahead, cp.async), so the FP64 warps issue no staging instructions at all.
//
// Phase 2 is slab_phase2 of apply_slab.cuh unchanged (same arithmetic, bit-identical quadrature-point values).
#pragma once
#include "apply_slab.cuh"

namespace jots {

constexpr int PATCH_INVALID = (int)0x80000000;   // node slot of a patch that no element of the patch touches

template <int D, int PX, int PY>
struct PatchDims {
    static constexpr int NXN = (D - 1) * PX + 1, NYN = (D - 1) * PY + 1;
    static constexpr int NPEN = NXN * NYN;            // unique nodal pencils
    static constexpr int NN = NPEN * D;               // nodes, [k][J][I]
    static constexpr int E = PX * PY;
    static constexpr int REC = (NN + E + 3) / 4 * 4;  // ints per descriptor: nodes, element ids, padding (16-byte multiple)
};

// plane stride (one quadrature plane of the whole patch): the four quadrature planes that the lanes of a half-warp read
// at the column offsets {0, D-1, 2(D-1)} must fall into distinct 8-byte banks
constexpr int patch_plane_stride(int npen) {
    int ps = npen;
    while (ps % 16 != 5 && ps % 16 != 11) ++ps;
    return ps;
}

// Work items of phase 3: (pencil, up to two (element, local ij) visits); pencils shared by four elements are split in two
// items that meet in the atomic.  item = pencil | nv << 8 | (el << 5 | ij) << 10 | (el << 5 | ij) << 20
struct PatchItems {
    unsigned item[128];
    int n;
};

template <int D, int PX, int PY>
inline PatchItems build_patch_items() {
    typedef PatchDims<D, PX, PY> PD;
    PatchItems it;
    it.n = 0;
    for (int J = 0; J < PD::NYN; ++J)
        for (int I = 0; I < PD::NXN; ++I) {
            int xs[2][2], ys[2][2], nx = 0, ny = 0;   // (element coordinate, local index)
            if (I % (D - 1)) { xs[nx][0] = I / (D - 1); xs[nx][1] = I % (D - 1); ++nx; }
            else {
                if (I / (D - 1) < PX) { xs[nx][0] = I / (D - 1); xs[nx][1] = 0; ++nx; }
                if (I > 0) { xs[nx][0] = I / (D - 1) - 1; xs[nx][1] = D - 1; ++nx; }
            }
            if (J % (D - 1)) { ys[ny][0] = J / (D - 1); ys[ny][1] = J % (D - 1); ++ny; }
            else {
                if (J / (D - 1) < PY) { ys[ny][0] = J / (D - 1); ys[ny][1] = 0; ++ny; }
                if (J > 0) { ys[ny][0] = J / (D - 1) - 1; ys[ny][1] = D - 1; ++ny; }
            }
            const unsigned pencil = (unsigned)(J * PD::NXN + I);
            auto visit = [&](int a, int b) { return (unsigned)(((ys[b][0] * PX + xs[a][0]) << 5) | (ys[b][1] * D + xs[a][1])); };
            if (nx * ny == 4) {
                for (int b = 0; b < 2; ++b) it.item[it.n++] = pencil | 2u << 8 | visit(0, b) << 10 | visit(1, b) << 20;
            } else if (nx * ny == 2) {
                it.item[it.n++] = pencil | 2u << 8 | visit(0, 0) << 10 | (nx == 2 ? visit(1, 0) : visit(0, 1)) << 20;
            } else it.item[it.n++] = pencil | 1u << 8 | visit(0, 0) << 10;
        }
    return it;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}

template <int D, int Q, int NP, int PX, int PY>
struct PatchSmem {
    typedef PatchDims<D, PX, PY> PD;
    static constexpr int PS = patch_plane_stride(PD::NPEN), PSET = Q * PS;
    static constexpr int ELS = 2 * Q * D * D;                       // result records of one element: [qz][h][D*D]
    static constexpr int planes = NP * PSET + (NP * PSET) % 2;      // doubles
    static constexpr int out = 2 * PD::E * ELS;                     // A then B
    static constexpr int stage = 2 * 2 * PD::NN + (2 * 2 * PD::NN) % 2;   // [buffer][x | z][node]
    static constexpr size_t total = (size_t)(planes + out + stage) * sizeof(double) + (size_t)3 * PD::REC * sizeof(int);
};

template <int D, int Q, int OP, int CF, int PX, int PY, int T>
__global__ void __launch_bounds__(T) form_apply_patch_kernel(const FormArgs a, const Slab2Tables<D, Q> t, const PatchItems items) {
    if (a.skip && *a.skip != 0.0) return;   // enqueued ahead of a Krylov solve that has already stopped
    typedef SlabCfg<OP, CF> C;
    typedef PatchDims<D, PX, PY> PD;
    typedef PatchSmem<D, Q, C::NP, PX, PY> SM;
    constexpr int NF = C::NF, NFA = NF > 0 ? NF : 1;
    constexpr int NPEN = PD::NPEN, NN = PD::NN, E = PD::E, REC = PD::REC, RS = PD::NXN;
    constexpr int PS = SM::PS, PSET = SM::PSET, ELS = SM::ELS, DD = D * D;
    constexpr bool MIRROR = Slab3Layout<D, Q>::MIRROR;
    constexpr bool NEED_Z = C::Z_GRAD || NF > 0;
    constexpr int NWORK = T - 32;                       // threads of the FP64 warps in phases 1 / 3
    static_assert(T % 32 == 0 && T >= E * 2 * Q, "phase 2 needs 2Q threads per element");
    static_assert(NPEN <= NWORK, "phase 1 is one pass of the worker warps");
    extern __shared__ __align__(16) double smem[];
    double* planes = smem;
    double* outA = smem + SM::planes;
    double* outB = outA + E * ELS;
    double* stage = outB + E * ELS;
    int* ring = reinterpret_cast<int*>(stage + SM::stage);            // [3][REC]
    const int tid = threadIdx.x, lane = tid & 31;
    const bool copier = tid >= NWORK;                   // warp-uniform
    const int ntiles = a.n_patches;
    const int nmine = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const unsigned my_item = tid < items.n ? items.item[tid] : 0u;

    // ---- the copy warp ----------------------------------------------------------------------------------------------
    auto issue_desc = [&](int i) {
        if (i < nmine) {
            const int* src = a.patches + (size_t)((int)blockIdx.x + i * (int)gridDim.x) * REC;
            int* dst = ring + (i % 3) * REC;
#pragma unroll
            for (int m = 0; m < (REC / 4 + 31) / 32; ++m) {
                const int c = lane + 32 * m;
                if (c < REC / 4) cp_async16(dst + 4 * c, src + 4 * c);
            }
        }
    };
    auto issue_vals = [&](int i) {
        if (i < nmine) {
            const int* dsc = ring + (i % 3) * REC;
            double* sx = stage + (size_t)(i % 2) * 2 * NN;
            double* sz = sx + NN;
            int g[(NN + 31) / 32];
#pragma unroll
            for (int m = 0; m < (NN + 31) / 32; ++m) { const int n = lane + 32 * m; g[m] = n < NN ? dsc[n] : PATCH_INVALID; }
#pragma unroll
            for (int m = 0; m < (NN + 31) / 32; ++m) {
                const int n = lane + 32 * m;
                if (n < NN) {
                    if (g[m] == PATCH_INVALID) { sx[n] = 0.0; if (NEED_Z) sz[n] = 0.0; }
                    else {
                        const int idx = g[m] < 0 ? ~g[m] : g[m];
                        if (g[m] < 0 && a.mask_in) sx[n] = 0.0;     // eliminated column
                        else cp_async8(sx + n, a.x + idx);
                        if (NEED_Z) cp_async8(sz + n, a.z + idx);
                    }
                }
            }
        }
    };

    if (copier) { issue_desc(0); issue_desc(1); cp_async_commit(); cp_async_wait<0>(); }
    __syncthreads();
    if (copier) { issue_vals(0); cp_async_commit(); cp_async_wait<0>(); }
    __syncthreads();

    double gxx = 0, gyy = 0, gzz = 0, det = 0;
    if (a.geom_stride == 0) {
        const double* gm = a.geom;
        gxx = gm[0] * gm[0] + gm[1] * gm[1] + gm[2] * gm[2];
        gyy = gm[3] * gm[3] + gm[4] * gm[4] + gm[5] * gm[5];
        gzz = gm[6] * gm[6] + gm[7] * gm[7] + gm[8] * gm[8];
        det = gm[9];
    }

    for (int it = 0; it < nmine; ++it) {
        const int* dsc = ring + (it % 3) * REC;
        // ---------------- phase 1: one worker thread per unique pencil | copies for the next tiles ------------------------
        if (copier) {
            issue_desc(it + 2);
            issue_vals(it + 1);
            cp_async_commit();
        } else if (tid < NPEN) {
            const double* sx = stage + (size_t)(it % 2) * 2 * NN;
            const double* sz = sx + NN;
            double xr[D], zr[D], fr[NFA][D];
#pragma unroll
            for (int k = 0; k < D; ++k) {
                xr[k] = sx[k * NPEN + tid];
                if (NEED_Z) {
                    zr[k] = sz[k * NPEN + tid];
                    if (NF > 0) {
                        double f[NFA];
                        nodal_fields<CF, OP>(a.mat, zr[k], f);
#pragma unroll
                        for (int n = 0; n < NF; ++n)
                            fr[n][k] = (CF == CF_K || CF == CF_KL) ? (a.ck * a.inv_rc) * f[n] : (CF == CF_LINFIELDS && n == 2) ? a.ck * f[n] : f[n];
                    }
                }
            }
            double* base = planes + tid;
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

        // ---------------- phase 2: half a quadrature plane per thread (all warps) ----------------------------------------
        if (tid < E * Q * 2) {
            const int el = tid / (2 * Q), r = tid - el * 2 * Q, qz = r >> 1;
            const bool h = r & 1;
            const int eid = dsc[NN + el];
            if (eid >= 0) {
                if (a.geom_stride != 0) {
                    const double* gm = a.geom + (size_t)eid * a.geom_stride;
                    gxx = gm[0] * gm[0] + gm[1] * gm[1] + gm[2] * gm[2];
                    gyy = gm[3] * gm[3] + gm[4] * gm[4] + gm[5] * gm[5];
                    gzz = gm[6] * gm[6] + gm[7] * gm[7] + gm[8] * gm[8];
                    det = gm[9];
                }
                const double wz = t.W[qz] * det;
                const int ex = el % PX, ey = el / PX;
                double* pl = planes + qz * PS + (D - 1) * ey * RS + (D - 1) * ex;
                slab_phase2<D, Q, OP, CF, PSET, MIRROR, RS, false>(a, t, pl, h, wz, gxx, gyy, gzz, 0u, outA + el * ELS + r * DD,
                                                                     outB + el * ELS + r * DD);
            }
        }
        __syncthreads();

        // ---------------- phase 3: transposed z contraction + scatter, one item per worker thread ------------------------
        if (copier) cp_async_wait<0>();      // values of tile it+1 and the descriptor of tile it+2 have landed
        else if (tid < items.n) {
            const int pencil = my_item & 0xff, nv = (my_item >> 8) & 3;
            double Aq[Q], Bq[Q];
#pragma unroll
            for (int qz = 0; qz < Q; ++qz) { Aq[qz] = 0.0; Bq[qz] = 0.0; }
            bool any = false;
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const unsigned ev = (my_item >> (10 + 10 * v)) & 0x3ffu;
                const int el = ev >> 5, ij = ev & 31;
                if (v < nv && dsc[NN + el] >= 0) {
                    any = true;
                    const double* oa = outA + el * ELS + ij;
                    const double* ob = outB + el * ELS + ij;
#pragma unroll
                    for (int qz = 0; qz < Q; ++qz) {
                        Aq[qz] += oa[(2 * qz) * DD] + oa[(2 * qz + 1) * DD];
                        Bq[qz] += ob[(2 * qz) * DD] + ob[(2 * qz + 1) * DD];
                    }
                }
            }
            if (any) {
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    double s = 0.0;
#pragma unroll
                    for (int qz = 0; qz < Q; ++qz) s += t.B[qz * D + k] * Aq[qz] + t.G[qz * D + k] * Bq[qz];
                    const int g = dsc[k * NPEN + pencil];
                    if (g >= 0) atomicAdd(a.y + g, s);
                    else if (g != PATCH_INVALID) {
                        if (!a.mask_out) atomicAdd(a.y + (~g), s);
                        else if (a.diag_one) a.y[~g] = a.x[~g];   // every writer stores the same value
                    }
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace jots
