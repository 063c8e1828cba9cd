// Instantiations of the register-slab kernel (hexes, diagonal metric, CF_CONST / CF_K).
#include <cstdlib>
#include <mutex>

#include "apply_slab.cuh"
#include "diag_face_host.hpp"
#include "slab_tables.hpp"
#include "launch.hpp"

namespace jots {

// cudaFuncSetAttribute / occupancy are per device: configure every kernel once per device (several
// contexts of one process may sit on different GPUs, one host thread each)
struct PerDevice {
    std::mutex m;
    int grid_max[64] = {0};
    template <class K>
    int get(K kern, int threads, size_t smem) {
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lock(m);
        if (!grid_max[dev]) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            int sms = 148, per_sm = 1;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
            grid_max[dev] = sms * (per_sm > 0 ? per_sm : 1);
        }
        return grid_max[dev];
    }
};

template <int D, int Q, int OP, int CF, int E>
static int launch_slab3e(const FormArgs& a, const TablesRT& r, cudaStream_t st) {
    // thread count follows phase 2 (2Q threads per element); phases 1/3 read shared memory and loop
    constexpr int T = (E * 2 * Q + 31) / 32 * 32;
    typedef SlabCfg<OP, CF> C;
    Slab2Tables<D, Q> t;
    fill_slab2_tables<D, Q>(r, t);
    const size_t smem = Slab3Smem<D, Q, C::NP, E>::total;
    auto kern = form_apply_slab3_kernel<D, Q, OP, CF, E, T>;
    static PerDevice pd;
    const int grid_max = pd.get(kern, T, smem);
    const int ntiles = (a.nelem + E - 1) / E;
    static const std::string label = kernel_label("form_apply_slab3_kernel", D, Q, OP, CF, ("," + std::to_string(E) + "," + std::to_string(T)).c_str());
    static const std::string label_dot = label.substr(0, label.size() - 1) + ",dot>", label_side = label.substr(0, label.size() - 1) + ",dot,side>";
    if ((a.dot || a.side_chunk) && (Slab3Stages<Q, E>::value != 2 || OP == OP_RES || !a.dot)) return 1;   // see slab_kernel_dot_tiles
    last_element_kernel() = a.side_chunk ? label_side.c_str() : a.dot ? label_dot.c_str() : label.c_str();
    kern<<<ntiles < grid_max ? ntiles : grid_max, T, smem, st>>>(a, t);
    return 0;
}
// Q3 tile: 6 elements on 64 threads (four blocks of two warps per SM) for the bilinear forms -- the blocks' latency-bound phases 1
// and 3 overlap other blocks' phase 2 better than with 12 elements on 128 threads (two blocks per SM): Jacobian apply 1.072 ->
// 1.054 ms, with inner product and side stream 1.104 -> 1.089 ms; the residual form is faster with 12 (0.977 vs 0.987 ms)
// (profiles/r3m).  JOTS_SLAB3_Q3_E = 6 | 12 forces one size for every form.
static int q3_tile(int op) {
    static const int forced = [] { const char* e = getenv("JOTS_SLAB3_Q3_E"); return e ? atoi(e) : 0; }();
    if (forced == 6 || forced == 12) return forced;
    return op == OP_RES ? 12 : 6;
}

template <int D, int Q, int OP, int CF>
static int launch_slab3(const FormArgs& a, const TablesRT& r, cudaStream_t st) {
    if constexpr (Q == 4) {
        // Q2: 16-element tiles with one staging buffer keep four blocks per SM and all 128 threads busy in phase 2 (measured
        // against E = 12 / 14 / 16 with two buffers, profiles/README.md r1j); the rho(T) / C(T) forms need the smaller tile
        if constexpr (CF != CF_FULL && CF != CF_FULLC) return launch_slab3e<D, Q, OP, CF, 16>(a, r, st);
        else return launch_slab3e<D, Q, OP, CF, 14>(a, r, st);
    }
    else if constexpr (Q == 5) return q3_tile(OP) == 6 ? launch_slab3e<D, Q, OP, CF, 6>(a, r, st) : launch_slab3e<D, Q, OP, CF, 12>(a, r, st);
    else return launch_slab3e<D, Q, OP, CF, 32>(a, r, st);
}

static bool slab_klin() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("JOTS_SLAB_KLIN"); v = e ? atoi(e) : 1; }
    return v != 0;
}

template <int D, int Q>
static int dispatch(int op, int cf, const FormArgs& a, const TablesRT& t, cudaStream_t st) {
    // k(T) linear (two coefficients, the reference's own cases): k'_h is uniform, one property field less in the
    // Jacobian apply (JOTS_SLAB_KLIN=0 keeps the general kernel)
    if (op == OP_JAC && cf == CF_K && a.mat.k.n == 2 && slab_klin()) return launch_slab3<D, Q, OP_JAC, CF_KL>(a, t, st);
#define CASE(OPV, CFV) if (op == OPV && cf == CFV) return launch_slab3<D, Q, OPV, CFV>(a, t, st);
    CASE(OP_LIN, CF_CONST) CASE(OP_LIN, CF_K) CASE(OP_JAC, CF_K) CASE(OP_RES, CF_CONST) CASE(OP_RES, CF_K)
#undef CASE
    // rho(T) / C(T): pipelined kernel only
    if (op == OP_LIN && cf == CF_LINFIELDS) return launch_slab3<D, Q, OP_LIN, CF_LINFIELDS>(a, t, st);
    if constexpr (Q < 5) {
        const bool rho_uniform = a.mat.rho.n == 1;   // five property fields instead of eight
        if (op == OP_JAC && cf == CF_FULL)
            return rho_uniform ? launch_slab3<D, Q, OP_JAC, CF_FULLC>(a, t, st) : launch_slab3<D, Q, OP_JAC, CF_FULL>(a, t, st);
        if (op == OP_RES && cf == CF_FULL)
            return rho_uniform ? launch_slab3<D, Q, OP_RES, CF_FULLC>(a, t, st) : launch_slab3<D, Q, OP_RES, CF_FULL>(a, t, st);
    }
    return 1;
}

template <int D, int Q, int OP, int CF>
static int launch_diag_slab(const FormArgs& a, const TablesRT& r, cudaStream_t st) {
    constexpr int E = (Q == 4) ? 14 : (Q == 5) ? 12 : 32;      // Q3: 83 KB of planes per block (OP_JAC, CF_K), two blocks per SM
    constexpr int T = ((E * (2 * Q > D * D ? 2 * Q : D * D)) + 31) / 32 * 32;
    constexpr int NF = CoefNF<CF, OP>::value;
    constexpr int NPI = NF + (OP == OP_JAC ? 2 : 0);
    SlabDiagTables<D, Q> t;
    for (int i = 0; i < D; ++i) t.mB[i] = 0.0;
    for (int q = 0; q < Q; ++q) {
        t.W[q] = r.W[q];
        const double half = (Q % 2 == 1 && q == Q / 2) ? 0.5 : 1.0;
        for (int i = 0; i < D; ++i) {
            const double b = r.B[q * D + i], g = r.G[q * D + i], w = r.W[q];
            t.B[q * D + i] = b; t.G[q * D + i] = g;
            t.B2[q * D + i] = b * b; t.G2[q * D + i] = g * g; t.BG[q * D + i] = b * g;
            t.B2W[q * D + i] = w * b * b; t.G2W[q * D + i] = w * g * g; t.BGW[q * D + i] = w * b * g;
            t.B2Wh[q * D + i] = half * w * b * b; t.G2Wh[q * D + i] = half * w * g * g; t.BGWh[q * D + i] = half * w * b * g;
            t.mB[i] += w * b * b;
        }
    }
    const size_t smem = (size_t)E * Elem2Stride<D, Q, NPI + 2>::value * sizeof(double);
    auto kern = form_diag_slab2_kernel<D, Q, OP, CF, E, T>;
    static PerDevice pd;
    pd.get(kern, T, smem);
    kern<<<(a.nelem + E - 1) / E, T, smem, st>>>(a, t);
    return 0;
}

template <int D, int Q>
static int dispatch_diag(int op, int cf, const FormArgs& a, const TablesRT& t, cudaStream_t st) {
#define CASE(OPV, CFV) if (op == OPV && cf == CFV) return launch_diag_slab<D, Q, OPV, CFV>(a, t, st);
    CASE(OP_LIN, CF_CONST) CASE(OP_LIN, CF_K) CASE(OP_JAC, CF_K) CASE(OP_LIN, CF_LINFIELDS)
#undef CASE
    return 1;
}

int launch_form_diag_slab(int p, int op, int cf, const FormArgs& a, const TablesRT& t, cudaStream_t st) {
    if (p == 1) return dispatch_diag<2, 3>(op, cf, a, t, st);
    if (p == 2) return dispatch_diag<3, 4>(op, cf, a, t, st);
    if (p == 3) return dispatch_diag<4, 5>(op, cf, a, t, st);
    return 1;
}

// Tiles of the gather-table kernel launch_form_apply_slab would run for this form when that kernel honours FormArgs::dot (two
// staging buffers: every order but the 16-element Q2 tiles; bilinear forms only), else 0.  *side_capacity = entries per tile the
// side stream (FormArgs::side_*) can carry: two pairs per thread.
template <int D, int Q>
static int dot_tiles(int op, int cf, const FormArgs& a, int* side_capacity) {
    if (op == OP_RES) return 0;
    int E = (Q == 5) ? q3_tile(op) : (Q == 4) ? 16 : 32;
    if (Q == 4 && cf == CF_FULL) E = 14;
    const bool full = cf == CF_FULL;
    const bool covered = (op == OP_LIN && (cf == CF_CONST || cf == CF_K || cf == CF_LINFIELDS)) || (op == OP_JAC && (cf == CF_K || (full && Q < 5)));
    if (!covered || (Q == 4 && E >= 16)) return 0;
    const int T = (E * 2 * Q + 31) / 32 * 32;
    if (side_capacity) *side_capacity = 4 * T;
    return (a.nelem + E - 1) / E;
}
int slab_kernel_dot_tiles(int p, int op, int cf, const FormArgs& a, int* side_capacity) {
    if (a.nelem <= 0) return 0;
    if (p == 1) return dot_tiles<2, 3>(op, cf, a, side_capacity);
    if (p == 2) return dot_tiles<3, 4>(op, cf, a, side_capacity);
    if (p == 3) return dot_tiles<4, 5>(op, cf, a, side_capacity);
    return 0;
}

// returns 0 when launched, non-zero when this (order, op, cf) has no slab instantiation
int launch_form_apply_slab(int p, int op, int cf, const FormArgs& a, const TablesRT& t, cudaStream_t st) {
    if (p == 1) return dispatch<2, 3>(op, cf, a, t, st);
    if (p == 2) return dispatch<3, 4>(op, cf, a, t, st);
    if (p == 3) return dispatch<4, 5>(op, cf, a, t, st);
    return 1;
}

}  // namespace jots
