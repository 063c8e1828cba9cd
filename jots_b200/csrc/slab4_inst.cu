// Instantiations of the row-structured slab kernel (apply_slab4.cuh): hexes with a diagonal metric whose elements
// form contiguous rows (refined bricks).
#include <cstdlib>
#include <mutex>

#include "apply_slab4.cuh"
#include "diag_face_host.hpp"
#include "slab_tables.hpp"
#include "launch.hpp"

namespace jots {

namespace {
struct PerDevice4 {
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

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
}  // namespace

// elements per tile of the row-structured kernel for order p (0: no instantiation); the host builds its RowTile
// table for exactly this length
int slab4_tile_elems(int p) {
    if (p == 1) return 32;
    if (p == 2) return env_int("JOTS_SLAB4_ELEMS", 16) == 14 ? 14 : 16;   // read when a context is built
    if (p == 3) return 12;
    return 0;
}

template <int D, int Q, int OP, int CF, int E, bool TMA_ROWS>
static int launch_slab4s(const FormArgs& a, const TablesRT& r, cudaStream_t st) {
    constexpr int T = (E * 2 * Q + 31) / 32 * 32;
    typedef SlabCfg<OP, CF> C;
    constexpr bool NEED_Z = C::Z_GRAD || C::NF > 0;
    Slab2Tables<D, Q> t;
    fill_slab2_tables<D, Q>(r, t);
    const size_t smem = Slab4Smem<D, Q, C::NP, E, NEED_Z ? 2 : 1>::total;
    auto kern = form_apply_slab4_kernel<D, Q, OP, CF, E, T, TMA_ROWS>;
    static PerDevice4 pd;
    const int grid_max = pd.get(kern, T, smem);
    const int ntiles = a.n_rowtiles;
    kern<<<ntiles < grid_max ? ntiles : grid_max, T, smem, st>>>(a, t);
    last_element_kernel() = "form_apply_slab4_kernel";
    return 0;
}

// JOTS_SLAB4_STAGE=cpasync: rows by 8-byte cp.async instead of TMA bulk copies (A/B; Q2 only)
template <int D, int Q, int OP, int CF, int E>
static int launch_slab4e(const FormArgs& a, const TablesRT& r, cudaStream_t st) {
    if constexpr (Q == 4 && E == 16) {
        static const bool cpasync = [] { const char* e = getenv("JOTS_SLAB4_STAGE"); return e && e[0] == 'c'; }();
        if (cpasync) return launch_slab4s<D, Q, OP, CF, E, false>(a, r, st);
    }
    return launch_slab4s<D, Q, OP, CF, E, true>(a, r, st);
}

template <int D, int Q, int OP, int CF>
static int launch_slab4(const FormArgs& a, const TablesRT& r, cudaStream_t st) {
    if constexpr (Q == 4) {
        if (a.rowtile_elems == 14) return launch_slab4e<D, Q, OP, CF, 14>(a, r, st);
        if (a.rowtile_elems == 16) return launch_slab4e<D, Q, OP, CF, 16>(a, r, st);
    } else if constexpr (Q == 5) {
        if (a.rowtile_elems == 12) return launch_slab4e<D, Q, OP, CF, 12>(a, r, st);
    } else if (a.rowtile_elems == 32) return launch_slab4e<D, Q, OP, CF, 32>(a, r, st);
    return 1;
}

template <int D, int Q>
static int dispatch4(int op, int cf, const FormArgs& a, const TablesRT& t, cudaStream_t st) {
#define CASE(OPV, CFV) if (op == OPV && cf == CFV) return launch_slab4<D, Q, OPV, CFV>(a, t, st);
    CASE(OP_LIN, CF_CONST) CASE(OP_LIN, CF_K) CASE(OP_JAC, CF_K) CASE(OP_RES, CF_CONST) CASE(OP_RES, CF_K)
    CASE(OP_LIN, CF_LINFIELDS)
#undef CASE
    if constexpr (Q < 5) {
        const bool rho_uniform = a.mat.rho.n == 1;   // five property fields instead of eight
        if (op == OP_JAC && cf == CF_FULL)
            return rho_uniform ? launch_slab4<D, Q, OP_JAC, CF_FULLC>(a, t, st) : launch_slab4<D, Q, OP_JAC, CF_FULL>(a, t, st);
        if (op == OP_RES && cf == CF_FULL)
            return rho_uniform ? launch_slab4<D, Q, OP_RES, CF_FULLC>(a, t, st) : launch_slab4<D, Q, OP_RES, CF_FULL>(a, t, st);
    }
    return 1;
}

// returns 0 when launched, non-zero when this (order, op, cf) has no instantiation or the space has no row tiles
int launch_form_apply_slab4(int p, int op, int cf, const FormArgs& a, const TablesRT& t, cudaStream_t st) {
    if (!a.rowtiles || a.n_rowtiles <= 0) return 1;
    if (p == 1) return dispatch4<2, 3>(op, cf, a, t, st);
    if (p == 2) return dispatch4<3, 4>(op, cf, a, t, st);
    if (p == 3) return dispatch4<4, 5>(op, cf, a, t, st);
    return 1;
}

}  // namespace jots
