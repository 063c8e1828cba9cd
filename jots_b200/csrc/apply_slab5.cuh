// EXPERIMENTAL (opt-in, JOTS_SLAB_VARIANT=6) -- written at the end of round 1 after the GPU budget was spent: it
// compiles for sm_100a but HAS NOT RUN ON HARDWARE YET; its parity tests (tests/test_gpu_slab5.py) are skipped unless
// JOTS_TEST_EXPERIMENTAL=1.  Nothing selects it by default.
//
// Warp-specialised version of the pipelined gather-table slab kernel (apply_slab.cuh, v3) for Q2 tiles of 16 elements.
//
// In v3 every warp walks through phase 1 -> barrier -> phase 2 -> barrier -> phase 3 -> barrier, and the profile
// (profiles/README.md, r1m) shows what that costs: phase 2, the FP64-dense part (56 % of the instructions), is only
// 29 % of the warp time, 20 % of the warp time waits at barriers, and on average 1.5 of the 4 warps of a scheduler are
// in phase 2 -- hence an FP64 pipe at 57 %.  Here a block is two warp groups with fixed roles:
//   * warps 0-3, the "plane" warps, run phase 2 (slab_phase2) back to back, tile after tile, on double-buffered planes;
//   * warps 4-7, the "pencil" warps, stage the gather rows and L-vector values (cp.async), run phase 1 of tile i+1 and
//     phase 3 of tile i, i.e. everything that is latency-bound.
// Hand-over is by two mbarrier pairs, no block-wide barrier after start-up:
//   full[b]  pencil -> plane : the planes of buffer b hold the phase-1 output of a tile
//   done[b]  plane -> pencil : phase 2 has replaced them by its result planes
// A pencil thread owns its plane slots (element, ij) in both phases, so the reuse of buffer b for tile i+2 after phase 3
// of tile i is ordered by program order within the thread; the plane warps cannot start tile i+2 before full[b]
// completes again, which is after that phase 3.  Two blocks of 256 threads per SM at 128 registers: 8 plane warps that
// never leave the FP64-dense code (v3: 16 warps x 29 %).
#pragma once
#include "apply_slab.cuh"
#include "apply_slab4.cuh"   // mbarrier helpers

namespace jots {

template <int D, int Q, int NP, int E> struct Slab5Smem {
    static constexpr int ND = D * D * D;
    static constexpr size_t planes = (size_t)2 * E * Elem3Stride<D, Q, NP>::value * sizeof(double);   // two buffers
    static constexpr size_t stage = (size_t)2 * E * ND * sizeof(double);                              // x and z of one tile
    static constexpr size_t idx = (size_t)4 * E * ND * sizeof(int);                                   // four tiles in flight
    static constexpr size_t bars = 4 * sizeof(unsigned long long);
    static constexpr size_t total = planes + stage + idx + bars;
};

__device__ __forceinline__ void pencil_group_barrier() { asm volatile("bar.sync 1, 128;\n" ::: "memory"); }

// SPLIT: the E*D*D - 128 left-over pencils of phases 1 and 3 as one more pass of short pieces (as in apply_slab.cuh,
// where the pieces are validated) instead of a second full pass on 16 lanes -- the pencil group is the latency-critical
// side of the hand-over.
template <int D, int Q, int OP, int CF, int E, bool SPLIT = false>
__global__ void __launch_bounds__(256, 2) form_apply_slab5_kernel(const FormArgs a, const Slab2Tables<D, Q> t) {
    typedef SlabCfg<OP, CF> C;
    constexpr int NF = C::NF, NFA = NF > 0 ? NF : 1, NP = C::NP;
    constexpr int PS = Elem3Stride<D, Q, NP>::PS, PSET = Elem3Stride<D, Q, NP>::PSET;
    constexpr int ES = Elem3Stride<D, Q, NP>::value;
    constexpr bool MIRROR = Slab3Layout<D, Q>::MIRROR;
    constexpr int ND = D * D * D, DD = D * D;
    constexpr bool NEED_Z = C::Z_GRAD || NF > 0;
    constexpr int TP = 128;                                  // threads per warp group
    static_assert(E * 2 * Q == TP, "one plane thread per (element, quadrature plane, half)");
    constexpr int LEFT = E * D * D - TP;
    static_assert(!SPLIT || (LEFT > 0 && LEFT * 2 * Q <= TP && Q % 2 == 0), "split pass: left-over pieces must fit one pass");
    __shared__ double s_tab[SPLIT ? 2 * Q * D : 1];         // B then G, indexed at run time by the split pass
    if (SPLIT && threadIdx.x == 0) {
#pragma unroll
        for (int m = 0; m < Q * D; ++m) { s_tab[m] = t.B[m]; s_tab[Q * D + m] = t.G[m]; }
    }
    extern __shared__ __align__(16) double smem[];
    double* planes = smem;                                                  // [2][E*ES]
    double* sx = smem + (size_t)2 * E * ES;                                 // [E*ND]
    double* sz = sx + E * ND;                                               // [E*ND]
    int* idxbuf = reinterpret_cast<int*>(sz + E * ND);                      // [4][E*ND]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(idxbuf + 4 * E * ND);   // [2]
    unsigned long long* done = full + 2;                                    // [2]
    const int tid = threadIdx.x;
    const int ntiles = (a.nelem + E - 1) / E;
    const int nmine = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    auto tile_of = [&](int i) { return (int)blockIdx.x + i * (int)gridDim.x; };
    auto nvalid = [&](int i) { const int r = a.nelem - tile_of(i) * E; return r < E ? r : E; };

    if (tid == 0) {
        mbar_init(&full[0], TP);
        mbar_init(&full[1], TP);
        mbar_init(&done[0], TP);
        mbar_init(&done[1], TP);
        mbar_fence_init();
    }
    __syncthreads();   // the only block-wide barrier

    if (tid < TP) {
        // =============================== plane warps: phase 2, tile after tile =================================
        const int el = tid / (2 * Q), r = tid - el * 2 * Q, qz = r >> 1;
        const bool h = r & 1;
        double gxx = 0, gyy = 0, gzz = 0, det = 0;
        if (a.geom_stride == 0) {
            const double* gm = a.geom;
            gxx = gm[0] * gm[0] + gm[1] * gm[1] + gm[2] * gm[2];
            gyy = gm[3] * gm[3] + gm[4] * gm[4] + gm[5] * gm[5];
            gzz = gm[6] * gm[6] + gm[7] * gm[7] + gm[8] * gm[8];
            det = gm[9];
        }
#pragma unroll 1
        for (int it = 0; it < nmine; ++it) {
            if (a.geom_stride != 0) {
                int e = tile_of(it) * E + el;
                if (e >= a.nelem) e = a.nelem - 1;
                const double* gm = a.geom + (size_t)e * a.geom_stride;
                gxx = gm[0] * gm[0] + gm[1] * gm[1] + gm[2] * gm[2];
                gyy = gm[3] * gm[3] + gm[4] * gm[4] + gm[5] * gm[5];
                gzz = gm[6] * gm[6] + gm[7] * gm[7] + gm[8] * gm[8];
                det = gm[9];
            }
            const double wz = t.W[qz] * det;
            double* pl = planes + (size_t)(it & 1) * E * ES + el * ES + qz * PS;
            mbar_wait(&full[it & 1], (unsigned)((it >> 1) & 1));
            slab_phase2<D, Q, OP, CF, PSET, MIRROR>(a, t, pl, h, wz, gxx, gyy, gzz, 0xffffffffu);
            mbar_arrive(&done[it & 1]);
        }
    } else {
        // =============================== pencil warps: staging, phase 1, phase 3 ===============================
        const int ptid = tid - TP;
        // gather rows of tile i -> idxbuf[i % 4]
        auto issue_idx = [&](int i) {
            if (i < nmine) {
                const int n = nvalid(i) * ND;
                const int* src = a.gather + (size_t)tile_of(i) * E * ND;
                int* dst = idxbuf + (i & 3) * E * ND;
#pragma unroll 1
                for (int m = ptid; m < n; m += TP) cp_async4(dst + m, src + m);
            }
            cp_async_commit();
        };
        // L-vector values of tile i (its gather rows are visible in shared memory) -> the staging buffer
        auto issue_vals = [&](int i) {
            if (i < nmine) {
                const int n = nvalid(i) * ND;
                const int* ix = idxbuf + (i & 3) * E * ND;
#pragma unroll 1
                for (int m = ptid; m < n; m += TP) {
                    const int g = ix[m];
                    const int idx = g < 0 ? ~g : g;
                    cp_async8(sx + m, a.x + idx);
                    if (NEED_Z) cp_async8(sz + m, a.z + idx);
                }
            }
            cp_async_commit();
        };
        // phase 1 of tile i: z contraction of the staged values into the planes of buffer i % 2
        auto phase1 = [&](int i) {
            const int* ix = idxbuf + (i & 3) * E * ND;
            double* pb = planes + (size_t)(i & 1) * E * ES;
            const int nit = nvalid(i) * DD;
#pragma unroll 1
            for (int item = ptid; item < (SPLIT && nit > TP ? TP : nit); item += TP) {
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
                double* base = pb + el * ES + ij;
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
            if constexpr (SPLIT) {
                // left-over pencils: piece = (pencil, quadrature plane qz, x-planes | z-planes and property fields)
                const int item = TP + ptid / (2 * Q), part = ptid - (ptid / (2 * Q)) * (2 * Q), qz = part >> 1;
                if (item < nit) {
                    const int el = item / DD, ij = item - el * DD;
                    double* base = pb + el * ES + ij;
                    const double* tb = s_tab + qz * D;
                    const double* tg = s_tab + Q * D + qz * D;
                    if ((part & 1) == 0) {
                        double sv = 0.0, sg = 0.0;
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            const int m = el * ND + ij + DD * k;
                            const double xv = (ix[m] < 0 && a.mask_in) ? 0.0 : sx[m];
                            sv += tb[k] * xv;
                            sg += tg[k] * xv;
                        }
                        base[C::P_XB * PSET + qz * PS] = sv;
                        if (C::X_GRAD) base[C::P_XG * PSET + qz * PS] = sg;
                    } else if (NEED_Z) {
                        double zs = 0.0, zg = 0.0, fs[NFA];
#pragma unroll
                        for (int n = 0; n < NFA; ++n) fs[n] = 0.0;
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            const double zv = sz[el * ND + ij + DD * k];
                            zs += tb[k] * zv;
                            zg += tg[k] * zv;
                            if (NF > 0) {
                                double f[NFA];
                                nodal_fields<CF, OP>(a.mat, zv, f);
#pragma unroll
                                for (int n = 0; n < NF; ++n)
                                    fs[n] += tb[k] * ((CF == CF_K || CF == CF_KL) ? (a.ck * a.inv_rc) * f[n] : (CF == CF_LINFIELDS && n == 2) ? a.ck * f[n] : f[n]);
                            }
                        }
                        if (C::Z_GRAD) { base[C::P_ZB * PSET + qz * PS] = zs; base[C::P_ZG * PSET + qz * PS] = zg; }
#pragma unroll
                        for (int n = 0; n < NF; ++n) base[(C::P_F0 + n) * PSET + qz * PS] = fs[n];
                    }
                }
            }
        };
        // phase 3 of tile i: transposed z contraction of the result planes of buffer i % 2 + scatter
        auto phase3 = [&](int i) {
            const int* ix = idxbuf + (i & 3) * E * ND;
            const double* pb = planes + (size_t)(i & 1) * E * ES;
            const int nit = nvalid(i) * DD;
#pragma unroll 1
            for (int item = ptid; item < (SPLIT && nit > TP ? TP : nit); item += TP) {
                const int el = item / DD, ij = item - el * DD;
                const double* base = pb + el * ES + ij;
                double Aq[Q], Bq[Q];
#pragma unroll
                for (int qz = 0; qz < Q; ++qz) {
                    Aq[qz] = base[qz * PS] + base[PSET + qz * PS];
                    Bq[qz] = base[2 * PSET + qz * PS] + base[3 * PSET + qz * PS];
                }
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    double sa = 0.0, sb = 0.0;
#pragma unroll
                    for (int qz = 0; qz < Q; ++qz) { sa += t.B[qz * D + k] * Aq[qz]; sb += t.G[qz * D + k] * Bq[qz]; }
                    const double s = sa + sb;
                    const int g = ix[el * ND + ij + DD * k];
                    if (g >= 0) atomicAdd(a.y + g, s);
                    else if (!a.mask_out) atomicAdd(a.y + (~g), s);
                    else if (a.diag_one) a.y[~g] = a.x[~g];   // every writer stores the same value
                }
            }
            if constexpr (SPLIT) {
                // left-over pencils: piece = (pencil, output node k, lower | upper half of the quadrature planes)
                const int item = TP + ptid / (2 * D), r = ptid - (ptid / (2 * D)) * (2 * D), k = r >> 1, q0 = (r & 1) * (Q / 2);
                if (ptid < LEFT * 2 * D && item < nit) {
                    const int el = item / DD, ij = item - el * DD;
                    const double* base = pb + el * ES + ij;
                    double s = 0.0;
#pragma unroll
                    for (int dq = 0; dq < Q / 2; ++dq) {
                        const int qz = q0 + dq;
                        s += s_tab[qz * D + k] * (base[qz * PS] + base[PSET + qz * PS]) +
                             s_tab[Q * D + qz * D + k] * (base[2 * PSET + qz * PS] + base[3 * PSET + qz * PS]);
                    }
                    const int g = ix[el * ND + ij + DD * k];
                    if (g >= 0) atomicAdd(a.y + g, s);
                    else if (!a.mask_out) atomicAdd(a.y + (~g), s);
                    else if (a.diag_one) a.y[~g] = a.x[~g];   // every writer stores the same value
                }
            }
        };

        // start-up: gather rows of tiles 0..3, values of tile 0, phase 1 of tile 0, values of tile 1
        issue_idx(0);
        issue_idx(1);
        issue_idx(2);
        issue_idx(3);
        cp_async_wait<0>();
        pencil_group_barrier();
        issue_vals(0);
        cp_async_wait<0>();
        pencil_group_barrier();
        if (nmine > 0) {
            phase1(0);
            mbar_arrive(&full[0]);
        }
        pencil_group_barrier();              // the staging buffer has been read by everybody
        issue_vals(1);
        cp_async_wait<0>();
        pencil_group_barrier();
        // invariant at the top of iteration it: the values of tile it+1 and the gather rows of tiles it+1, it+2 (and it+3)
        // are visible in shared memory; slot (it+4) % 4 still holds the rows of tile it
#pragma unroll 1
        for (int it = 0; it < nmine; ++it) {
            if (it + 1 < nmine) {
                phase1(it + 1);                              // buffer (it+1) % 2: its last reader was phase 3 of tile it-1 (this thread)
                mbar_arrive(&full[(it + 1) & 1]);
            }
            pencil_group_barrier();                          // the staging buffer is free
            issue_vals(it + 2);
            mbar_wait(&done[it & 1], (unsigned)((it >> 1) & 1));
            phase3(it);
            cp_async_wait<0>();                              // values of tile it+2 and the rows issued one iteration ago
            pencil_group_barrier();                          // ... visible to the group; nobody reads the rows of tile it any more
            issue_idx(it + 4);
        }
        cp_async_wait<0>();
    }
}

}  // namespace jots
