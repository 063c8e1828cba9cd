// Row-structured variant of the pipelined slab kernel (apply_slab.cuh, v3) for meshes whose elements come in runs
// along the fastest lattice axis with contiguous L-vector numbering (refined bricks: every BASELINE synthetic case).
//
// v3 reads an explicit gather table (4 (p+1)^3 / p^3 bytes per DOF, 13.5 B/DOF at Q2 -- more than half of the
// compulsory 24 B/DOF of a Jacobian apply) and stages the L-vector values with one 8-byte cp.async per node.  Here a
// tile of E elements is described by D^2 row starts (RowTile, 128 bytes for Q2 instead of 1728 bytes of gather
// rows): node (i, j, k) of element el is base[k*D + j] + p*el + i, so the D^2 rows of p*E + 1 contiguous doubles are
// brought into shared memory by 1-D TMA bulk copies (cp.async.bulk, completion on an mbarrier) -- 2 D^2 copies per
// tile for x and z instead of 2 E D^3 scattered ones, and nodes shared by neighbouring elements of the row are
// read once.  cp.async.bulk wants 16-byte aligned source, destination and size; a row may start on an odd double,
// so the copy starts one double early and the consumer adds the shift (the first / last row of a vector falls back
// to plain loads instead of reading outside the vector).
//
// Phase 2 and the shared-memory plane layout are those of v3 (the same function, slab_phase2).  Phases 1 and 3 run
// over the *unique* nodal pencils of the tile -- (p E + 1) D of them instead of E D^2, a node shared by two elements
// of the row is contracted once and its planes are written to / summed from both elements -- which removes a third
// of their arithmetic and of the FP64 atomics, makes both fit one pass of the block at E = 16 (99 of 128 threads,
// so phase 2 can use all 128), and turns the scatter into contiguous runs: lane n adds to base[k*D + j] + n.  No
// index ever travels through memory.
//
// A pencil thread owns its plane slots (element, ij) in both phases, so phase 3 of tile i and phase 1 of tile i+1
// need no barrier between them: they are fused into one pass ("phase 3+1"), which leaves two block barriers per
// tile and lets the latency-bound tails of one overlap the loads of the other.  Pencils and the issue of the bulk
// copies are spread evenly over the warps (a bulk copy is a warp-serialised uniform-datapath instruction).
#pragma once
#include "apply_slab.cuh"

namespace jots {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    const unsigned b = smem_u32(bar);
    unsigned done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(b), "r"(parity)
            : "memory");
    } while (!done);
}
// 1-D TMA bulk copy global -> shared, completion (bytes) signalled on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

template <int D, int Q, int NP, int E, int NV>
struct Slab4Smem {
    static constexpr int P = D - 1, DD = D * D;
    static constexpr int LMAX = P * E + 1;                                // nodes of a full row
    static constexpr int RS0 = (LMAX + 2) / 2 * 2;                        // + alignment shift, rounded to even
    static constexpr int RS = RS0 + ((6 - RS0 % 16) + 16) % 16;           // == 6 (mod 16), even
    static constexpr int NROW = NV * DD;                                  // rows per stage (x rows, then z rows)
    static constexpr size_t planes = (size_t)E * Elem3Stride<D, Q, NP>::value * sizeof(double);
    static constexpr size_t rows = (size_t)2 * NROW * RS * sizeof(double);
    static constexpr size_t desc = (size_t)3 * sizeof(RowTile<D>);
    static constexpr size_t bars = 5 * sizeof(unsigned long long);
    static constexpr size_t total = planes + rows + desc + bars;
    static_assert(LMAX <= 64, "one 64-bit essential mask per row");
    static_assert(NROW <= 32, "the rows of a tile are issued by one warp");
    static_assert(planes % 16 == 0 && rows % 16 == 0 && sizeof(RowTile<D>) % 16 == 0, "bulk copies need 16-byte alignment");
};

// Q2 with up to two property fields: cap at 128 registers so that four blocks stay resident per SM
template <int Q, int CF> struct Slab4MinBlocks { static constexpr int value = Q != 4 ? 1 : CF == CF_FULL ? 2 : CF == CF_FULLC ? 3 : 4; };

// TMA_ROWS: the rows come by cp.async.bulk (one warp-serialised UBLKCP per row); otherwise by coalesced 8-byte cp.async
// issued by all threads (no alignment window, completion through cp.async.mbarrier.arrive on the same mbarrier).  The
// tile descriptor always comes by one bulk copy.
template <int D, int Q, int OP, int CF, int E, int T, bool TMA_ROWS>
__global__ void __launch_bounds__(T, Slab4MinBlocks<Q, CF>::value) form_apply_slab4_kernel(const FormArgs a, const Slab2Tables<D, Q> t) {
    if (a.skip && *a.skip != 0.0) return;   // enqueued ahead of a Krylov solve that has already stopped
    typedef SlabCfg<OP, CF> C;
    constexpr int NF = C::NF, NFA = NF > 0 ? NF : 1, NP = C::NP;
    constexpr int PS = Elem3Stride<D, Q, NP>::PS, PSET = Elem3Stride<D, Q, NP>::PSET;
    constexpr int ES = Elem3Stride<D, Q, NP>::value;
    constexpr bool MIRROR = Slab3Layout<D, Q>::MIRROR;
    constexpr int DD = D * D, P = D - 1;
    constexpr bool NEED_Z = C::Z_GRAD || NF > 0;
    constexpr int NV = NEED_Z ? 2 : 1;
    typedef Slab4Smem<D, Q, NP, E, NV> S;
    constexpr int RS = S::RS, NROW = S::NROW, NNMAX = S::LMAX;
    constexpr int NW = T / 32;
    constexpr int NPEN = NNMAX * D;                        // unique pencils of a full tile
    constexpr int PPW = (NPEN + NW - 1) / NW;              // pencils per warp when one pass suffices
    constexpr bool ONE_PASS = PPW <= 32;
    constexpr int RPW = (NROW + NW - 1) / NW;              // bulk copies issued per warp (by its last RPW lanes)
    extern __shared__ __align__(16) double smem[];
    double* planes = smem;
    double* rows = smem + (size_t)E * ES;                                        // [2][NROW][RS]
    RowTile<D>* desc = reinterpret_cast<RowTile<D>*>(rows + (size_t)2 * NROW * RS);   // [3]
    unsigned long long* rowbar = reinterpret_cast<unsigned long long*>(desc + 3);     // [2]
    unsigned long long* descbar = rowbar + 2;                                    // [3]
    const RowTile<D>* gdesc = static_cast<const RowTile<D>*>(a.rowtiles);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = a.n_rowtiles;
    const int nmine = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    auto tile_of = [&](int i) { return (int)blockIdx.x + i * (int)gridDim.x; };
    // a row that starts on an odd double (16-byte phase 8) is copied from one double earlier
    const int xpar = (int)((reinterpret_cast<unsigned long long>(a.x) >> 3) & 1ull);
    const int zpar = NEED_Z ? (int)((reinterpret_cast<unsigned long long>(a.z) >> 3) & 1ull) : 0;

    if (tid == 0) {
        mbar_init(&rowbar[0], TMA_ROWS ? NROW : T);
        mbar_init(&rowbar[1], TMA_ROWS ? NROW : T);
        mbar_init(&descbar[0], 1);
        mbar_init(&descbar[1], 1);
        mbar_init(&descbar[2], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // descriptor of tile i -> desc[i % 3]
    auto issue_desc = [&](int i) {
        if (tid == 0 && i < nmine) {
            mbar_arrive_expect_tx(&descbar[i % 3], (unsigned)sizeof(RowTile<D>));
            bulk_g2s(&desc[i % 3], gdesc + tile_of(i), (unsigned)sizeof(RowTile<D>), &descbar[i % 3]);
        }
    };
    // L-vector rows of tile i -> rows[i % 2]; row r (x rows first, then z rows) is issued by lane 32 - RPW + r % RPW of
    // warp r / RPW
    auto issue_rows = [&](int i) {
        if constexpr (!TMA_ROWS) {
            if (i < nmine) {
                mbar_wait(&descbar[i % 3], (unsigned)((i / 3) & 1));
                const RowTile<D>& d = desc[i % 3];
                const int L = P * d.nv + 1;
                double* stage = rows + (size_t)(i % 2) * NROW * RS;
#pragma unroll 1
                for (int c = tid; c < NROW * NNMAX; c += T) {
                    const int r = c / NNMAX, m = c - r * NNMAX;
                    const int v = r / DD, rr = r - v * DD;
                    if (m < L) cp_async8(stage + r * RS + m, (v ? a.z : a.x) + ((long long)d.base[rr] + m));
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&rowbar[i % 2])) : "memory");
            }
            return;
        }
        const int r = warp * RPW + (lane - (32 - RPW));
        if (lane >= 32 - RPW && r < NROW && i < nmine) {
            mbar_wait(&descbar[i % 3], (unsigned)((i / 3) & 1));
            const RowTile<D>& d = desc[i % 3];
            const int v = r / DD, rr = r - v * DD;
            const int L = P * d.nv + 1;
            const int base = d.base[rr];
            const int sh = (base + (v ? zpar : xpar)) & 1;
            const int cnt = (L + sh + 1) & ~1;
            const double* src = (v ? a.z : a.x) + base;
            double* dst = rows + ((size_t)(i % 2) * NROW + r) * RS;
            void* bar = &rowbar[i % 2];
            if (base - sh >= 0 && (long long)(base - sh + cnt) <= a.nvec) {
                mbar_arrive_expect_tx(bar, (unsigned)cnt * 8u);
                bulk_g2s(dst, src - sh, (unsigned)cnt * 8u, bar);
            } else {
                // first / last row of the vector: the aligned window would leave it
#pragma unroll 1
                for (int m = 0; m < L; ++m) dst[sh + m] = src[m];
                fence_proxy_async();   // order these generic-proxy writes before the bulk copy that reuses the slot
                mbar_arrive(bar);
            }
        }
    };

    // phase 1 of tile i for pencil (node n, j): z contraction of x, z and the nodal property fields -> planes
    auto phase1 = [&](int i, int n, int j) {
        const RowTile<D>& d = desc[i % 3];
        const int nv = d.nv;
        if (n >= P * nv + 1) return;
        const bool ess_tile = d.any_ess != 0;
        const double* rx = rows + (size_t)(i % 2) * NROW * RS;
        const double* rz = rx + DD * RS;
        double xr[D], zr[D], fr[NFA][D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const int rr = k * D + j;
            const int b = d.base[rr];
            xr[k] = rx[rr * RS + (TMA_ROWS ? ((b + xpar) & 1) : 0) + n];
            if (ess_tile && a.mask_in && ((d.ess[rr] >> n) & 1ull)) xr[k] = 0.0;
            if (NEED_Z) {
                zr[k] = rz[rr * RS + (TMA_ROWS ? ((b + zpar) & 1) : 0) + n];
                if (NF > 0) {
                    double f[NFA];
                    nodal_fields<CF, OP>(a.mat, zr[k], f);
#pragma unroll
                    for (int m = 0; m < NF; ++m)
                        fr[m][k] = CF == CF_K ? (a.ck * a.inv_rc) * f[m] : (CF == CF_LINFIELDS && m == 2) ? a.ck * f[m] : f[m];
                }
            }
        }
        // node n is local node ib of element ea (if that element exists) and, when it closes element ea - 1,
        // local node P of that one
        const int ea = n / P, ib = n - ea * P;
        const bool has0 = ea < nv, has1 = (ib == 0 && ea >= 1);
        double* d0 = planes + ea * ES + j * D + ib;
        double* d1 = d0 - ES + P;
#pragma unroll
        for (int qz = 0; qz < Q; ++qz) {
            double s = 0.0, sg = 0.0, zs = 0.0, zg = 0.0, fs[NFA];
#pragma unroll
            for (int k = 0; k < D; ++k) {
                s += t.B[qz * D + k] * xr[k];
                if (C::X_GRAD) sg += t.G[qz * D + k] * xr[k];
                if (C::Z_GRAD) { zs += t.B[qz * D + k] * zr[k]; zg += t.G[qz * D + k] * zr[k]; }
            }
#pragma unroll
            for (int m = 0; m < NF; ++m) {
                fs[m] = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) fs[m] += t.B[qz * D + k] * fr[m][k];
            }
            if (has0) {
                d0[C::P_XB * PSET + qz * PS] = s;
                if (C::X_GRAD) d0[C::P_XG * PSET + qz * PS] = sg;
                if (C::Z_GRAD) { d0[C::P_ZB * PSET + qz * PS] = zs; d0[C::P_ZG * PSET + qz * PS] = zg; }
#pragma unroll
                for (int m = 0; m < NF; ++m) d0[(C::P_F0 + m) * PSET + qz * PS] = fs[m];
            }
            if (has1) {
                d1[C::P_XB * PSET + qz * PS] = s;
                if (C::X_GRAD) d1[C::P_XG * PSET + qz * PS] = sg;
                if (C::Z_GRAD) { d1[C::P_ZB * PSET + qz * PS] = zs; d1[C::P_ZG * PSET + qz * PS] = zg; }
#pragma unroll
                for (int m = 0; m < NF; ++m) d1[(C::P_F0 + m) * PSET + qz * PS] = fs[m];
            }
        }
    };
    // phase 3 of tile i for pencil (node n, j): sum the result planes of the elements sharing the node, transposed z
    // contraction, contiguous scatter (lane n -> base[k*D + j] + n)
    auto phase3 = [&](int i, int n, int j) {
        const RowTile<D>& d = desc[i % 3];
        const int nv = d.nv;
        if (n >= P * nv + 1) return;
        const bool ess_tile = d.any_ess != 0;
        const int ea = n / P, ib = n - ea * P;
        const bool has0 = ea < nv, has1 = (ib == 0 && ea >= 1);
        const double* s0 = planes + ea * ES + j * D + ib;
        const double* s1 = s0 - ES + P;
        double Aq[Q], Bq[Q];
#pragma unroll
        for (int qz = 0; qz < Q; ++qz) { Aq[qz] = 0.0; Bq[qz] = 0.0; }
        if (has0) {
#pragma unroll
            for (int qz = 0; qz < Q; ++qz) {
                Aq[qz] = s0[qz * PS] + s0[PSET + qz * PS];
                Bq[qz] = s0[2 * PSET + qz * PS] + s0[3 * PSET + qz * PS];
            }
        }
        if (has1) {
#pragma unroll
            for (int qz = 0; qz < Q; ++qz) {
                Aq[qz] += s1[qz * PS] + s1[PSET + qz * PS];
                Bq[qz] += s1[2 * PSET + qz * PS] + s1[3 * PSET + qz * PS];
            }
        }
#pragma unroll
        for (int k = 0; k < D; ++k) {
            double sa = 0.0, sb = 0.0;   // two independent chains
#pragma unroll
            for (int qz = 0; qz < Q; ++qz) { sa += t.B[qz * D + k] * Aq[qz]; sb += t.G[qz * D + k] * Bq[qz]; }
            const double s = sa + sb;
            const int rr = k * D + j;
            const long long g = (long long)d.base[rr] + n;
            const bool is_ess = ess_tile && ((d.ess[rr] >> n) & 1ull);
            if (!is_ess || !a.mask_out) atomicAdd(a.y + g, s);
            else if (a.diag_one) a.y[g] = a.x[g];   // every writer stores the same value
        }
    };
    // fused pass over the pencils: phase 3 of tile i3 (if >= 0), then phase 1 of tile i1 (if < nmine).  A pencil
    // thread owns its plane slots, so the only ordering needed is program order within the thread.
    auto pencil_pass = [&](int i3, int i1) {
        const bool do1 = i1 < nmine;
        if (do1) {
            mbar_wait(&descbar[i1 % 3], (unsigned)((i1 / 3) & 1));
            mbar_wait(&rowbar[i1 % 2], (unsigned)((i1 / 2) & 1));
        }
        if (ONE_PASS) {
            const int u = warp * PPW + lane;
            if (lane < PPW && u < NPEN) {
                const int j = u / NNMAX, n = u - j * NNMAX;
                if (i3 >= 0) phase3(i3, n, j);
                if (do1) phase1(i1, n, j);
            }
        } else {
#pragma unroll 1
            for (int u = tid; u < NPEN; u += T) {
                const int j = u / NNMAX, n = u - j * NNMAX;
                if (i3 >= 0) phase3(i3, n, j);
                if (do1) phase1(i1, n, j);
            }
        }
    };

    issue_desc(0);
    issue_desc(1);
    issue_rows(0);
    issue_rows(1);

    // geometry of a uniform mesh is loop invariant
    double gxx = 0, gyy = 0, gzz = 0, det = 0;
    if (a.geom_stride == 0) {
        const double* gm = a.geom;
        gxx = gm[0] * gm[0] + gm[1] * gm[1] + gm[2] * gm[2];
        gyy = gm[3] * gm[3] + gm[4] * gm[4] + gm[5] * gm[5];
        gzz = gm[6] * gm[6] + gm[7] * gm[7] + gm[8] * gm[8];
        det = gm[9];
    }

    pencil_pass(-1, 0);
    __syncthreads();

    for (int it = 0; it < nmine; ++it) {
        // slot (it+2) % 3 held the descriptor of tile it-1, dead since the pass that closed the previous iteration
        issue_desc(it + 2);

        // ---------------- phase 2: half a quadrature plane per thread ----------------------------
        const unsigned p2mask = __ballot_sync(0xffffffffu, tid < E * Q * 2);
        if (tid < E * Q * 2) {
            const int el = tid / (2 * Q), r = tid - el * 2 * Q, qz = r >> 1;
            const bool h = r & 1;
            if (a.geom_stride != 0) {
                int e = desc[it % 3].e0 + el;
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

        // ---------------- phase 3 of this tile + phase 1 of the next one ---------------------------
        // stage it % 2 held the rows of tile it (read by the previous pass): free for tile it+2, which lands while
        // tile it+1 is in phase 2
        issue_rows(it + 2);
        pencil_pass(it, it + 1);
        __syncthreads();
    }
}

}  // namespace jots
