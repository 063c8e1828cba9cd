// Instantiations of the experimental warp-specialised slab kernel (apply_slab5.cuh; opt-in, JOTS_SLAB_VARIANT=6, not yet
// run on hardware -- see the header).
#include <cstdlib>
#include <mutex>

#include "apply_slab5.cuh"
#include "launch.hpp"
#include "slab_tables.hpp"

namespace jots {

namespace {
struct PerDevice5 {
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
}  // namespace

template <int OP, int CF, bool SPLIT>
static int launch_slab5s(const FormArgs& a, const TablesRT& r, cudaStream_t st) {
    constexpr int D = 3, Q = 4, E = 16;
    typedef SlabCfg<OP, CF> C;
    Slab2Tables<D, Q> t;
    fill_slab2_tables<D, Q>(r, t);
    const size_t smem = Slab5Smem<D, Q, C::NP, E>::total;
    auto kern = form_apply_slab5_kernel<D, Q, OP, CF, E, SPLIT>;
    static PerDevice5 pd;
    const int grid_max = pd.get(kern, 256, smem);
    const int ntiles = (a.nelem + E - 1) / E;
    kern<<<ntiles < grid_max ? ntiles : grid_max, 256, smem, st>>>(a, t);
    return 0;
}

// JOTS_SLAB5_SPLIT=1: left-over pencils as short pieces
template <int OP, int CF>
static int launch_slab5(const FormArgs& a, const TablesRT& r, cudaStream_t st) {
    static const bool split = [] { const char* e = getenv("JOTS_SLAB5_SPLIT"); return e && e[0] == '1'; }();
    return split ? launch_slab5s<OP, CF, true>(a, r, st) : launch_slab5s<OP, CF, false>(a, r, st);
}

// Q2 hexes with a diagonal metric; returns non-zero when (op, cf) has no instantiation
int launch_form_apply_slab5(int p, int op, int cf, const FormArgs& a, const TablesRT& t, cudaStream_t st) {
    if (p != 2) return 1;
    if (op == OP_JAC && cf == CF_K && a.mat.k.n == 2) return launch_slab5<OP_JAC, CF_KL>(a, t, st);
#define CASE(OPV, CFV) if (op == OPV && cf == CFV) return launch_slab5<OPV, CFV>(a, t, st);
    CASE(OP_LIN, CF_CONST) CASE(OP_LIN, CF_K) CASE(OP_JAC, CF_K) CASE(OP_RES, CF_CONST) CASE(OP_RES, CF_K)
    CASE(OP_LIN, CF_LINFIELDS)
#undef CASE
    return 1;
}

}  // namespace jots
