// Two-phase unique-pencil patch kernel for the forms whose coefficients need no interpolated property field: uniform
// properties (CF_CONST) and k(T) LINEAR in T with uniform rho, C (the reference's own k(T) cases and BASELINE configs[3]).
//
// For a linear k(T) = c0 T + c1 the interpolant of the nodal k values IS c0 z_h + c1 (interpolation is linear and the basis
// is a partition of unity), so the quadrature-point conductivity comes from the interpolated state that the Jacobian /
// residual need anyway, and no nodal polynomial is evaluated at all.  What is left of phase 1 of apply_patch.cuh -- the
// contraction of x and z along the pencil direction -- is done by the phase-2 threads themselves, straight from the staged
// nodal values (27 + 27 shared-memory loads and FP64 FMAs per plane, all lanes of a quadrature column reading the same
// word: broadcast, conflict free).  That costs 7 % more FP64 instructions and removes a whole latency-bound phase with its
// barrier (27 % of a block's time in the three-phase patch kernel, profiles/r2b):
//
//   per tile:   [ all warps: quadrature planes -> result planes ]  barrier  [ worker warps: z-transpose + atomics ]
//
// with ONE block barrier per tile: the result planes are double buffered (the mirrored thread pair first adds its two half
// results with one shuffle exchange, so a buffer is 11 KB), hence phase 3 of tile t overlaps phase 2 of tile t+1 of the warps
// that are done.  The last warp is the block's copy engine (descriptors two tiles ahead, nodal values one tile ahead,
// cp.async; it zero-fills eliminated columns and unused slots) and joins phase 2.
#pragma once
#include "apply_patch.cuh"

namespace jots {

// result planes: word (r, el, j, i) at r*QS + el*ES + j*JS + i, r = 2 qz + (0: in-plane sum A, 1: z-derivative part B);
// (ES, JS, QS) found by exhaustive search: phase-2 stores conflict free, phase-3 loads 1.67 wavefronts on average
template <int D, int PX, int PY> struct Patch2Out { static constexpr int ES = D * D + 2, JS = D, QS = PX * PY * (D * D + 2) + 2; };

template <int D, int PX, int PY>
inline PatchItems build_patch2_items() {
    typedef Patch2Out<D, PX, PY> O;
    PatchItems it = build_patch_items<D, PX, PY>();
    for (int n = 0; n < it.n; ++n) {
        unsigned v = it.item[n], out = v & 0x3ffu;
        for (int s = 0; s < 2; ++s) {
            const unsigned ev = (v >> (10 + 10 * s)) & 0x3ffu, el = ev >> 5, ij = ev & 31;
            out |= (el * O::ES + (ij / D) * O::JS + ij % D) << (10 + 10 * s);
        }
        it.item[n] = out;     // pencil | nv << 8 | offset0 << 10 | offset1 << 20, element = offset / ES
    }
    return it;
}

// LAG (the warp-decoupled pipeline, see the kernel): four stage buffers and eight descriptor slots instead of two and four
template <int D, int Q, int PX, int PY, bool LAG = false>
struct Patch2Smem {
    typedef PatchDims<D, PX, PY> PD;
    typedef Patch2Out<D, PX, PY> O;
    static constexpr int NSTG = LAG ? 4 : 2, NRING = LAG ? 8 : 4;
    static constexpr int stage = NSTG * 2 * PD::NN + (NSTG * 2 * PD::NN) % 2;   // [buffer][x | z][node]
    static constexpr int out = 2 * 2 * Q * O::QS;                          // [buffer][r][QS]
    static constexpr int tab = 2 * Q * D + (2 * Q * D) % 2;                // B rows then G rows
    static constexpr size_t total = (size_t)(stage + out + tab) * sizeof(double) + (size_t)NRING * PD::REC * sizeof(int);
};

// shared-memory barriers with split arrive / wait (mbarrier): arrive releases, wait acquires
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(b);
    unsigned done = 0;
    // the hardware suspends the warp up to the hint (ns) per try, so a warp that is ahead does not spin on issue slots
#pragma unroll 1
    for (int spin = 0; spin < (1 << 16); ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity), "r"(100000u) : "memory");
        if (done) return;
    }
    __trap();      // a protocol error must not hang the device
}

// KLIN: k(T) = c0 T + c1 (else uniform properties).  OP as in forms.hpp.
// Measured dead ends of the decoupled pipeline (profiles/README.md r2z): rotating the warp that sits out phase 3 so that all four
// carry the same load (8 % / 6 % slower for a 4- / 6-tile rotation), shorter FMA chains in phase 3 (no change: its latency is no
// longer exposed), three or five blocks per SM (168 registers without spills: 6 % slower; 96 registers: 30 % slower), the fused
// inner product reading its inputs from eight x stage buffers instead of L2 (no change), or handed to the copy warp (the pencil
// sums or products left in shared memory: 2 % / 10 % slower -- the copy warp is not idle, it paces the block).
// Measured dead ends (profiles/README.md r2d, r2e), so that they are not tried again: issuing phase 3 of tile t-1 branch-free
// inside phase 2 of tile t (registers: 3 % slower); moving the copy-engine role to another warp with every tile to even out
// the four scheduler partitions (5 % slower).
// DOT: also add (x, A x) over the rows written to *a.dot (FormArgs::dot; the conjugate-gradient loop's (d, A d)).
// Side stream (FormArgs::side_*, DOT and LAG): measured dead end -- pulling the tile's slices into L2 at the start of the tile
// (prefetch.global.L2) so that their loads after phase 2 hit L2: 0.9084 -> 0.9094 ms (profiles/r3h).
//
// LAG: the warp-decoupled pipeline.  With one block barrier per tile every warp waits for the slowest one once per tile (14 % of
// the warp samples, profiles/r2v_phases_*).  Here phase 3 of tile t-1 runs AFTER phase 2 of tile t, and the hand-overs are
// mbarriers whose other side finished about one tile earlier, so a wait only blocks a warp that is a whole tile ahead:
//   READY[t & 3]  (copy warp -> all)   nodal values + descriptor of tile t have landed (fetched two / three tiles ahead)
//   FULL[t & 1]   (all -> all)         result planes of tile t written  (awaited before phase 3 of t, i.e. after phase 2 of t+1)
//   EMPTY[t & 1]  (all -> all)         phase 3 of tile t done           (awaited before the result stores of tile t+2)
// The copy warp's refill of stage buffer (t+2) & 3 follows its own FULL wait of tile t-2's successor, two tiles after the last read.
template <int D, int Q, int OP, bool KLIN, int PX, int PY, int T, bool DOT = false, bool LAG = false>
__global__ void __launch_bounds__(T, 4) form_apply_patch2_kernel(const FormArgs a, const Slab2Tables<D, Q> t, const PatchItems items) {
    if (a.skip && *a.skip != 0.0) return;   // enqueued ahead of a Krylov solve that has already stopped
    typedef PatchDims<D, PX, PY> PD;
    typedef Patch2Out<D, PX, PY> O;
    typedef Patch2Smem<D, Q, PX, PY, LAG> SM;
    constexpr int SMASK = SM::NSTG - 1, RMASK = SM::NRING - 1;
    constexpr int NPEN = PD::NPEN, NN = PD::NN, E = PD::E, REC = PD::REC, RS = PD::NXN;
    constexpr int DD = D * D, QC = (Q + 1) / 2, NQH = Q * QC;
    constexpr bool X_GRAD = (OP != OP_RES), Z_GRAD = (OP != OP_LIN), NEED_Z = Z_GRAD || KLIN;
    static_assert(T == E * 2 * Q, "one thread per (element, quadrature plane, half)");
    static_assert(Q % 2 == 0, "even quadrature rules only (no shared middle column)");
    extern __shared__ __align__(16) double smem[];
    double* stage = smem;
    double* outbuf = smem + SM::stage;
    double* s_tab = outbuf + SM::out;
    int* ring = reinterpret_cast<int*>(s_tab + SM::tab);                // [4][REC]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = T / 32;
    const bool copier = warp == NWARP - 1;              // the block's copy engine
    const unsigned my_item = tid < items.n ? items.item[tid] : 0u;
    const int ntiles = a.n_patches;
    if (tid < Q * D) { s_tab[tid] = t.B[tid]; s_tab[Q * D + tid] = t.G[tid]; }
    // Dynamic tile scheduling: the copy warp claims the block's i-th tile from a device counter two tiles ahead (s_tile[i & 3],
    // -1 = no tile left), so an SM that runs slower simply takes fewer tiles (a static round-robin left most SMs idle for the
    // last 8 % of the launch: sm__cycles_active avg 1.81 M against max 1.97 M, profiles/r2c).  The last block to leave resets
    // the counter for the next launch.
    __shared__ int s_tile[SM::NRING];
    __shared__ __align__(8) unsigned long long s_bar[8];     // LAG: READY[4], FULL[2], EMPTY[2]
    unsigned long long* const READY = s_bar; unsigned long long* const FULL = s_bar + 4; unsigned long long* const EMPTY = s_bar + 6;
    __shared__ double s_dot[DOT ? T : 1];    // fused (x, A x): one running sum per phase-3 thread
    if (DOT) s_dot[tid] = 0.0;
    int pending = 0;               // lane 0 of the copy warp: the claim made one tile earlier (its latency is never waited for)
    auto claim = [&](int i) {      // copy warp only
        int tile = 0;
        if (lane == 0) {
            tile = pending < ntiles ? pending : -1;
            pending = (int)atomicAdd(a.sched, 1u);
            s_tile[i & RMASK] = tile;
        }
        return __shfl_sync(0xffffffffu, tile, 0);
    };

    auto issue_desc = [&](int i, int tile) {
        if (tile >= 0) {
            const int* src = a.patches + (size_t)tile * REC;
            int* dst = ring + (i & RMASK) * REC;
#pragma unroll
            for (int m = 0; m < (REC / 4 + 31) / 32; ++m) {
                const int c = lane + 32 * m;
                if (c < REC / 4) cp_async16(dst + 4 * c, src + 4 * c);
            }
        }
    };
    auto issue_vals = [&](int i) {
        if (s_tile[i & RMASK] >= 0) {
            const int* dsc = ring + (i & RMASK) * REC;
            double* sx = stage + (size_t)(i & SMASK) * 2 * NN;
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

    if (LAG) {
        if (tid == 0) {
            for (int b = 0; b < 4; ++b) mbar_init(READY + b, 32);
            for (int b = 0; b < 2; ++b) { mbar_init(FULL + b, T); mbar_init(EMPTY + b, T); }
        }
        __syncthreads();
        if (copier) {
            if (lane == 0) pending = (int)atomicAdd(a.sched, 1u);
            const int t0 = claim(0), t1 = claim(1), t2 = claim(2);
            issue_desc(0, t0); issue_desc(1, t1); issue_desc(2, t2);
            cp_async_commit(); cp_async_wait<0>();
            __syncwarp();
            issue_vals(0); cp_async_commit(); cp_async_wait<0>(); mbar_arrive(READY + 0);
            issue_vals(1); cp_async_commit(); cp_async_wait<0>(); mbar_arrive(READY + 1);
        }
    } else {
        if (copier) {
            if (lane == 0) pending = (int)atomicAdd(a.sched, 1u);
            const int t0 = claim(0), t1 = claim(1);
            issue_desc(0, t0); issue_desc(1, t1);
            cp_async_commit(); cp_async_wait<0>();
        }
        __syncthreads();
        if (copier) { issue_vals(0); cp_async_commit(); cp_async_wait<0>(); }
        __syncthreads();
    }

    // thread constants of phase 2
    const int el = tid / (2 * Q), r = tid - el * 2 * Q, qz = r >> 1;
    const bool h = r & 1;
    const int ex = el % PX, ey = el / PX;
    const int node0 = (D - 1) * ey * RS + (D - 1) * ex;
    int mo[D];
#pragma unroll
    for (int i = 0; i < D; ++i) mo[i] = h ? D - 1 - i : i;
    double gxx = 0, gyy = 0, gzz = 0, det = 0;
    if (a.geom_stride == 0) {
        const double* gm = a.geom;
        gxx = gm[0] * gm[0] + gm[1] * gm[1] + gm[2] * gm[2];
        gyy = gm[3] * gm[3] + gm[4] * gm[4] + gm[5] * gm[5];
        gzz = gm[6] * gm[6] + gm[7] * gm[7] + gm[8] * gm[8];
        det = gm[9];
    }
    const SymTab<D, Q, 1> tB{t.B}, tBW{t.BW};
    const SymTab<D, Q, -1> tG{t.G}, tGW{t.GW};
    // linear conductivity: a1 = ck inv_rc (c0 z + c1), a2 = ck inv_rc c0;  uniform: a1 = ck a0
    const double s0 = KLIN ? a.ck * a.inv_rc * a.mat.k.c[0] : 0.0;
    const double s1 = KLIN ? a.ck * a.inv_rc * a.mat.k.c[1] : a.ck * a.a0;
    const double cmass = a.cm * a.m0;

    // phase 3 of tile i: transposed z contraction + scatter of this thread's item.  Branch free up to the atomics: the
    // loads are unconditional (every address is valid), contributions of holes are deselected.
    // sk (DOT): the pencil's output values, handed back for the fused inner product
    // xk (DOT, LAG): the pencil's own input values, re-read from the vector (L2) because the stage buffer may already be refilled
    auto phase3 = [&](int i, unsigned my_item, double (&sk)[D], double (&xk)[D]) {
        const int p3_pencil = my_item & 0xff, p3_nv = (my_item >> 8) & 3;
        const int p3_off0 = (my_item >> 10) & 0x3ff, p3_off1 = (my_item >> 20) & 0x3ff;
        const int* dsc = ring + (i & RMASK) * REC;
        const double* ob = outbuf + (size_t)(i & 1) * 2 * Q * O::QS;
        const bool v0 = dsc[NN + p3_off0 / O::ES] >= 0, v1 = p3_nv > 1 && dsc[NN + p3_off1 / O::ES] >= 0;
        int g[D];
#pragma unroll
        for (int k = 0; k < D; ++k) g[k] = dsc[k * NPEN + p3_pencil];
        if (DOT && LAG) {
#pragma unroll
            for (int k = 0; k < D; ++k)
                xk[k] = g[k] >= 0 ? a.x[g[k]] : ((g[k] != PATCH_INVALID && !a.mask_in) ? a.x[~g[k]] : 0.0);
        }
        double Aq[Q], Bq[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const double a0 = ob[(2 * q) * O::QS + p3_off0], a1v = ob[(2 * q) * O::QS + p3_off1];
            const double b0 = ob[(2 * q + 1) * O::QS + p3_off0], b1 = ob[(2 * q + 1) * O::QS + p3_off1];
            Aq[q] = (v0 ? a0 : 0.0) + (v1 ? a1v : 0.0);
            Bq[q] = (v0 ? b0 : 0.0) + (v1 ? b1 : 0.0);
        }
        const bool any = v0 || v1;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            // two short chains instead of one of length 2Q
            double sa = t.B[k] * Aq[0], sb = t.G[k] * Bq[0];
#pragma unroll
            for (int q = 1; q < Q; ++q) { sa += t.B[q * D + k] * Aq[q]; sb += t.G[q * D + k] * Bq[q]; }
            const double sum = sa + sb;
            if (DOT) sk[k] = sum;
            if (any) {
                if (g[k] >= 0) atomicAdd(a.y + g[k], sum);
                else if (g[k] != PATCH_INVALID) {
                    if (!a.mask_out) atomicAdd(a.y + (~g[k]), sum);
                    else if (a.diag_one) a.y[~g[k]] = a.x[~g[k]];   // every writer stores the same value
                }
            }
        }
    };

    // phase 3 of tile i for the threads that own an item, then the fused inner product
    auto phase3_all = [&](int i, double (&xk)[D]) {
        double sk[D];
        if (DOT) {
#pragma unroll
            for (int k = 0; k < D; ++k) { sk[k] = 0.0; if (LAG) xk[k] = 0.0; }
        }
        if (tid < items.n) phase3(i, my_item, sk, xk);
        if (DOT) {
            // Straight-line code after the divergent region, so that its latency overlaps the start of the next tile's phase 2:
            // with the FP64 pipe saturated by the other warps every dependent FP64 instruction inside phase 3 costs ~50 cycles
            // (measured: the same three FMAs inside the k loop +23 us per launch, here +16 us; a separate pass is 44 us).
            // No selection is needed: the staged input is zero wherever the row is not written (holes; eliminated columns --
            // the caller fuses only forms with mask_in == mask_out), and threads without an item keep sk = 0.
            double pr = xk[0] * sk[0];
#pragma unroll
            for (int k = 1; k < D; ++k) pr = fma(xk[k], sk[k], pr);
            s_dot[tid] += pr;
        }
    };

    const int side_chunk = (DOT && LAG) ? a.side_chunk : 0;
    const bool side_warp = side_chunk != 0 && !copier && 64 * warp < side_chunk;      // warp-uniform
    const int side_last = (int)a.nvec - 1;
    int it = 0;
    for (;; ++it) {
        if (LAG) mbar_wait(READY + (it & 3), (it >> 2) & 1);
        if (s_tile[it & RMASK] < 0) break;
        const int* dsc = ring + (it & RMASK) * REC;
        if (copier) {
            if (LAG) { issue_desc(it + 3, claim(it + 3)); issue_vals(it + 2); }
            else { issue_desc(it + 2, claim(it + 2)); issue_vals(it + 1); }
            cp_async_commit();
        }
        // ---------------- phase 2: quadrature planes from the staged nodal values, half a plane per thread ---------------
        {
            const double* sx = stage + (size_t)(it & SMASK) * 2 * NN + node0;
            const double* sz = sx + NN;
            const int eid = dsc[NN + el];
            if (a.geom_stride != 0) {
                const double* gm = a.geom + (size_t)(eid < 0 ? 0 : eid) * a.geom_stride;
                gxx = gm[0] * gm[0] + gm[1] * gm[1] + gm[2] * gm[2];
                gyy = gm[3] * gm[3] + gm[4] * gm[4] + gm[5] * gm[5];
                gzz = gm[6] * gm[6] + gm[7] * gm[7] + gm[8] * gm[8];
                det = gm[9];
            }
            const double wz = t.W[qz] * det;
            double tb[D], tg[D];
#pragma unroll
            for (int k = 0; k < D; ++k) { tb[k] = s_tab[qz * D + k]; tg[k] = s_tab[Q * D + qz * D + k]; }
            // plane of this quadrature layer: out[jj][i] = sum_k c[k] src[k][jj][mirrored i]
            auto zplane = [&](const double* src, const double (&c)[D], double (&out)[DD]) {
#pragma unroll
                for (int jj = 0; jj < D; ++jj)
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                        double s = c[0] * src[jj * RS + mo[i]];
#pragma unroll
                        for (int k = 1; k < D; ++k) s += c[k] * src[k * NPEN + jj * RS + mo[i]];
                        out[jj * D + i] = s;
                    }
            };
            double p9[DD], t6[D * QC], A[DD], Bp[DD];
#pragma unroll
            for (int i = 0; i < DD; ++i) { A[i] = 0.0; Bp[i] = 0.0; }
            double xB6[D * QC], xG6[D * QC], vy6[D * QC], zB6[D * QC], zG6[D * QC], a1[NQH], a2x[NQH];
            zplane(sx, tb, p9);
            xpassC<D, QC>(p9, tB, xB6);
            if (X_GRAD) xpassC<D, QC>(p9, tG, xG6);
            if (NEED_Z) {
                zplane(sz, tb, p9);
                xpassC<D, QC>(p9, tB, zB6);
                if (Z_GRAD) xpassC<D, QC>(p9, tG, zG6);
            }
            {
                double xv[NQH];
                ypassC<D, Q, QC>(xB6, tB, xv);                       // x at the quadrature points
                if (KLIN) {
                    ypassC<D, Q, QC>(zB6, tB, a1);                   // z at the quadrature points
#pragma unroll
                    for (int q = 0; q < NQH; ++q) a1[q] = s0 * a1[q] + s1;
                    if (OP == OP_JAC) {
#pragma unroll
                        for (int q = 0; q < NQH; ++q) a2x[q] = s0 * xv[q];
                    }
                }
                ypassTC<D, Q, QC>(xv, tBW, vy6);
            }
            const double c_m = cmass * wz;
            auto scaleF = [&](double (&F)[NQH], const double (&zd6)[D * QC], const SymTab<D, Q, 1>* tb_, const SymTab<D, Q, -1>* tg_) {
                if (KLIN) {
#pragma unroll
                    for (int q = 0; q < NQH; ++q) F[q] *= a1[q];
                    if (OP == OP_JAC) {
                        double zd[NQH];
                        if (tb_) ypassC<D, Q, QC>(zd6, *tb_, zd); else ypassC<D, Q, QC>(zd6, *tg_, zd);
#pragma unroll
                        for (int q = 0; q < NQH; ++q) F[q] += a2x[q] * zd[q];
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < NQH; ++q) F[q] *= s1;
                }
            };
            {   // x direction (both G's carry the mirror sign: it cancels)
                double F[NQH];
                if (OP == OP_RES) ypassC<D, Q, QC>(zG6, tB, F);
                else ypassC<D, Q, QC>(xG6, tB, F);
                scaleF(F, zG6, &tB, nullptr);
                ypassTC<D, Q, QC>(F, tBW, t6);
                const double cx = wz * gxx;
#pragma unroll
                for (int i = 0; i < D * QC; ++i) t6[i] *= cx;
                xpassTC_add<D, QC>(t6, tGW, A);
            }
            {   // y direction (+ mass)
                double F[NQH];
                if (OP == OP_RES) ypassC<D, Q, QC>(zB6, tG, F);
                else ypassC<D, Q, QC>(xB6, tG, F);
                scaleF(F, zB6, nullptr, &tG);
                ypassTC<D, Q, QC>(F, tGW, t6);
                const double cy = wz * gyy;
#pragma unroll
                for (int i = 0; i < D * QC; ++i) t6[i] = cy * t6[i] + c_m * vy6[i];
                xpassTC_add<D, QC>(t6, tBW, A);
            }
            {   // z direction: the G-contracted planes are formed here, not kept
                double F[NQH];
                zplane(OP == OP_RES ? sz : sx, tg, p9);
                xpassC<D, QC>(p9, tB, t6);
                ypassC<D, Q, QC>(t6, tB, F);
                if (KLIN && OP == OP_JAC) {
                    double zd6[D * QC];
                    zplane(sz, tg, p9);
                    xpassC<D, QC>(p9, tB, zd6);
                    scaleF(F, zd6, &tB, nullptr);
                } else scaleF(F, t6, &tB, nullptr);
                ypassTC<D, Q, QC>(F, tBW, t6);
                const double cz = wz * gzz;
#pragma unroll
                for (int i = 0; i < D * QC; ++i) t6[i] *= cz;
                xpassTC_add<D, QC>(t6, tBW, Bp);
            }
            flip_plane<D>(A, h);
            flip_plane<D>(Bp, h);
            // the pair (adjacent lanes) adds its two halves: h = 0 ends up with the in-plane sum, h = 1 with the z part
            double* op = outbuf + (size_t)(it & 1) * 2 * Q * O::QS + (size_t)r * O::QS + el * O::ES;
            if (LAG && it >= 2) mbar_wait(EMPTY + (it & 1), ((it - 2) >> 1) & 1);      // phase 3 of tile it-2 has read this buffer
#pragma unroll
            for (int m = 0; m < DD; ++m) {
                const double mine = h ? Bp[m] : A[m];
                const double other = __shfl_xor_sync(0xffffffffu, h ? A[m] : Bp[m], 1);
                op[(m / D) * O::JS + m % D] = mine + other;
            }
        }
        double xk[D];
        if (!LAG) {
            // fused (x, A x): the pencil's own input values are read from the stage buffer BEFORE the tile's barrier (the copy
            // warp refills that buffer right after it)
            if (DOT) {
                const double* sxp = stage + (size_t)(it & SMASK) * 2 * NN + (my_item & 0xff);
#pragma unroll
                for (int k = 0; k < D; ++k) xk[k] = sxp[k * NPEN];
            }
            if (copier) cp_async_wait<0>();      // nodal values of tile it+1 and the descriptor of tile it+2 have landed
            __syncthreads();
            // ---------------- phase 3: transposed z contraction + scatter (worker warps; the others run ahead) -----------
            phase3_all(it, xk);
        } else {
            mbar_arrive(FULL + (it & 1));
            if (copier) { cp_async_wait<0>(); __syncwarp(); mbar_arrive(READY + ((it + 2) & 3)); }
            // Side stream of the conjugate-gradient loop (FormArgs::side_*): this tile's slice of x += alpha d_old and of the
            // zeroing of the next output, one pair of entries per thread of the first warps (the copy warp paces the block and
            // takes no part; warps without entries skip the section).  The loads are issued here, where the register peak of
            // phase 2 is over, and consumed after phase 3 of the previous tile (about a microsecond later: their latency is not
            // waited for); streaming cache hints keep the slices from displacing the gathered nodal values in L2.  Vector
            // lengths are below 2^31 (DOF ids are 32-bit), so the index arithmetic is 32-bit.
            double2 side_xv = make_double2(0.0, 0.0), side_dv = side_xv;
            double side_al = 0.0;
            int side_k = -1;
            if (DOT && side_warp) {
                const int k = s_tile[it & RMASK] * side_chunk + 2 * tid;
                if (2 * tid < side_chunk && k < side_last) {          // a whole pair; an odd last entry is done after the loop
                    side_k = k;
                    side_xv = __ldcs(reinterpret_cast<const double2*>(a.side_x + k));
                    side_dv = __ldcs(reinterpret_cast<const double2*>(a.side_d + k));
                    side_al = __ldg(a.side_alpha);
                }
            }
            if (it >= 1) {
                mbar_wait(FULL + ((it - 1) & 1), ((it - 1) >> 1) & 1);
                phase3_all(it - 1, xk);
                mbar_arrive(EMPTY + ((it - 1) & 1));
            }
            if (DOT && side_k >= 0) {
                side_xv.x = fma(side_al, side_dv.x, side_xv.x);
                side_xv.y = fma(side_al, side_dv.y, side_xv.y);
                __stcs(reinterpret_cast<double2*>(a.side_x + side_k), side_xv);
                __stcs(reinterpret_cast<double2*>(a.side_zero + side_k), make_double2(0.0, 0.0));
            }
        }
    }
    if (LAG && it >= 1) {      // drain: phase 3 of the block's last tile
        double xk[D];
        mbar_wait(FULL + ((it - 1) & 1), ((it - 1) >> 1) & 1);
        phase3_all(it - 1, xk);
    }
    if (DOT && side_chunk != 0 && blockIdx.x == 0 && tid == 0 && (a.nvec & 1)) {      // the last entry of an odd-length vector
        a.side_x[side_last] = fma(__ldg(a.side_alpha), a.side_d[side_last], a.side_x[side_last]);
        a.side_zero[side_last] = 0.0;
    }
    if (DOT) {
        __syncthreads();
        if (warp == 0) {
            double v = 0.0;
#pragma unroll
            for (int m = 0; m < NWARP; ++m) v += s_dot[lane + 32 * m];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) atomicAdd(a.dot, v);
        }
    }
    // the last block to finish resets the tile counter (atomicInc wraps the exit counter to 0 by itself)
    if (tid == 0 && atomicInc(a.sched + 1, gridDim.x - 1) == gridDim.x - 1) a.sched[0] = 0u;
}

}  // namespace jots
