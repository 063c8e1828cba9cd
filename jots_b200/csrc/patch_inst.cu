// Instantiations of the unique-pencil patch kernel (apply_patch.cuh): Q2 hexes, 4 x 4 x 1 patches, 128 threads.
#include <cstdlib>
#include <mutex>

#include "apply_patch2.cuh"
#include "diag_face_host.hpp"
#include "launch.hpp"
#include "slab_tables.hpp"

namespace jots {

namespace {
struct PerDevicePatch {
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

template <int D, int Q, int OP, int CF, int PX, int PY>
int launch_patch(const FormArgs& a, const TablesRT& r, cudaStream_t st) {
    constexpr int T = (PX * PY * 2 * Q + 31) / 32 * 32;
    typedef SlabCfg<OP, CF> C;
    Slab2Tables<D, Q> t;
    fill_slab2_tables<D, Q>(r, t);
    static const PatchItems items = build_patch_items<D, PX, PY>();
    const size_t smem = PatchSmem<D, Q, C::NP, PX, PY>::total;
    auto kern = form_apply_patch_kernel<D, Q, OP, CF, PX, PY, T>;
    static PerDevicePatch pd;
    const int grid_max = pd.get(kern, T, smem);
    kern<<<a.n_patches < grid_max ? a.n_patches : grid_max, T, smem, st>>>(a, t, items);
    static const std::string label = kernel_label("form_apply_patch_kernel", D, Q, OP, CF, ("," + std::to_string(PX) + "," + std::to_string(PY) + "," + std::to_string(T)).c_str());
    last_element_kernel() = label.c_str();
    return 0;
}

template <int D, int Q, int OP, bool KLIN, int PX, int PY, bool DOT, bool LAG = false>
int launch_patch2_v(const FormArgs& a, const TablesRT& r, cudaStream_t st) {
    constexpr int T = PX * PY * 2 * Q;
    Slab2Tables<D, Q> t;
    fill_slab2_tables<D, Q>(r, t);
    static const PatchItems items = build_patch2_items<D, PX, PY>();
    const size_t smem = Patch2Smem<D, Q, PX, PY, LAG>::total;
    auto kern = form_apply_patch2_kernel<D, Q, OP, KLIN, PX, PY, T, DOT, LAG>;
    static PerDevicePatch pd;
    const int grid_max = pd.get(kern, T, smem);
    if (items.n > T - 32) return 1;      // phase 3 must be one pass of the worker warps
    kern<<<a.n_patches < grid_max ? a.n_patches : grid_max, T, smem, st>>>(a, t, items);
    static const std::string label = kernel_label("form_apply_patch2_kernel", D, Q, OP, KLIN ? CF_KL : CF_CONST,
                                                  ("," + std::to_string(PX) + "," + std::to_string(PY) + "," + std::to_string(T) + (DOT ? ",dot" : "") + (LAG ? ",lag" : "")).c_str());
    static const std::string label_side = label.substr(0, label.size() - 1) + ",side>";    // carrying the CG side stream
    last_element_kernel() = (DOT && LAG && a.side_chunk) ? label_side.c_str() : label.c_str();
    return 0;
}

// FormArgs::dot selects the variant that also forms (x, A x); the residual form is not bilinear and has none
// JOTS_PATCH2_LAG=0: one block barrier per tile instead of the warp-decoupled pipeline (A/B runs)
bool patch2_lag_enabled() {
    static const bool v = [] { const char* e = getenv("JOTS_PATCH2_LAG"); return !e || atoi(e) != 0; }();
    return v;
}

template <int D, int Q, int OP, bool KLIN, int PX, int PY>
int launch_patch2(const FormArgs& a, const TablesRT& r, cudaStream_t st) {
    const bool lag = patch2_lag_enabled();
    if (a.side_chunk && !(a.dot && lag)) return 1;      // the side stream exists in the dot variant of the decoupled pipeline only
    if (a.dot) {
        if (OP == OP_RES) return 1;
        constexpr int OPD = OP == OP_RES ? OP_LIN : OP;
        return lag ? launch_patch2_v<D, Q, OPD, KLIN, PX, PY, true, true>(a, r, st) : launch_patch2_v<D, Q, OPD, KLIN, PX, PY, true, false>(a, r, st);
    }
    return lag ? launch_patch2_v<D, Q, OP, KLIN, PX, PY, false, true>(a, r, st) : launch_patch2_v<D, Q, OP, KLIN, PX, PY, false, false>(a, r, st);
}

// JOTS_PATCH2=0 keeps the three-phase patch kernel for the forms the two-phase kernel covers (A/B runs)
bool patch2_enabled() {
    static const bool v = [] { const char* e = getenv("JOTS_PATCH2"); return !e || atoi(e) != 0; }();
    return v;
}

bool klin_enabled() {
    static const bool v = [] { const char* e = getenv("JOTS_SLAB_KLIN"); return !e || atoi(e) != 0; }();
    return v;
}

// the forms the two-phase kernel covers
bool patch2_covers(int op, int cf, const FormArgs& a) {
    if (!(patch2_enabled() && klin_enabled())) return false;
    const bool klin = cf == CF_K && a.mat.k.n == 2;
    if (!(cf == CF_CONST || klin)) return false;
    return op == OP_LIN || op == OP_RES || (op == OP_JAC && klin);
}

template <int D, int Q, int PX, int PY>
int dispatch_patch(int op, int cf, const FormArgs& a, const TablesRT& t, cudaStream_t st) {
    if (patch2_enabled() && klin_enabled()) {
        // uniform properties, or k(T) linear with uniform rho, C: no interpolated property field is needed
        const bool klin = cf == CF_K && a.mat.k.n == 2;
        if (cf == CF_CONST || klin) {
            if (op == OP_LIN) return klin ? launch_patch2<D, Q, OP_LIN, true, PX, PY>(a, t, st) : launch_patch2<D, Q, OP_LIN, false, PX, PY>(a, t, st);
            if (op == OP_JAC && klin) return launch_patch2<D, Q, OP_JAC, true, PX, PY>(a, t, st);
            if (op == OP_RES) return klin ? launch_patch2<D, Q, OP_RES, true, PX, PY>(a, t, st) : launch_patch2<D, Q, OP_RES, false, PX, PY>(a, t, st);
        }
    }
    if (op == OP_JAC && cf == CF_K && a.mat.k.n == 2 && klin_enabled()) return launch_patch<D, Q, OP_JAC, CF_KL, PX, PY>(a, t, st);
#define CASE(OPV, CFV) if (op == OPV && cf == CFV) return launch_patch<D, Q, OPV, CFV, PX, PY>(a, t, st);
    CASE(OP_LIN, CF_CONST) CASE(OP_LIN, CF_K) CASE(OP_JAC, CF_K) CASE(OP_RES, CF_CONST) CASE(OP_RES, CF_K)
    CASE(OP_LIN, CF_LINFIELDS)
#undef CASE
    const bool rho_uniform = a.mat.rho.n == 1;   // five property fields instead of eight
    if (op == OP_JAC && cf == CF_FULL)
        return rho_uniform ? launch_patch<D, Q, OP_JAC, CF_FULLC, PX, PY>(a, t, st) : launch_patch<D, Q, OP_JAC, CF_FULL, PX, PY>(a, t, st);
    if (op == OP_RES && cf == CF_FULL)
        return rho_uniform ? launch_patch<D, Q, OP_RES, CF_FULLC, PX, PY>(a, t, st) : launch_patch<D, Q, OP_RES, CF_FULL, PX, PY>(a, t, st);
    return 1;
}
}  // namespace

bool patch_kernel_fuses_dot(int p, int op, int cf, const FormArgs& a) {
    return p == 2 && op != OP_RES && a.patches && a.n_patches > 0 && patch2_covers(op, cf, a);
}

// entries of the CG side stream (FormArgs::side_*) one tile can carry: two per thread of the worker warps; 0 = not available
int patch_kernel_side_capacity(int p, int op, int cf, const FormArgs& a) {
    if (!patch_kernel_fuses_dot(p, op, cf, a) || !patch2_lag_enabled()) return 0;
    return 2 * (4 * 4 * 2 * 4 - 32);
}

void patch_shape(int p, int& px, int& py) {
    if (p == 2) { px = 4; py = 4; }
    else { px = 0; py = 0; }
}

// phase-3 work items of the kernel for order p (host copy, for the index-contract test); returns their number
int patch_items(int p, unsigned* out) {
    if (p != 2) return 0;
    const PatchItems it = build_patch_items<3, 4, 4>();   // the two-phase kernel re-encodes the same visits as offsets
    for (int i = 0; i < it.n; ++i) out[i] = it.item[i];
    return it.n;
}

// returns 0 when launched, non-zero when this (order, op, cf) has no patch instantiation
int launch_form_apply_patch(int p, int op, int cf, const FormArgs& a, const TablesRT& t, cudaStream_t st) {
    if (!a.patches || a.n_patches <= 0) return 1;
    if (p == 2) return dispatch_patch<3, 4, 4, 4>(op, cf, a, t, st);
    return 1;
}

}  // namespace jots
