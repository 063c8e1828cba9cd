#include "launch.hpp"

#include <cstdlib>

#include "diag_face.cuh"

namespace jots {

const char*& last_element_kernel() {
    static thread_local const char* name = "";
    return name;
}
std::string kernel_label(const char* base, int D, int Q, int op, int cf, const char* tail) {
    static const char* ops[] = {"OP_LIN", "OP_JAC", "OP_RES"};
    static const char* cfs[] = {"CF_CONST", "CF_LINFIELDS", "CF_K", "CF_FULL", "CF_FULLC", "CF_KL"};
    return std::string(base) + "<" + std::to_string(D) + "," + std::to_string(Q) + "," + ops[op] + "," + cfs[cf] + tail + ">";
}

int launch_form_apply(int dim, int p, int op, int cf, const FormArgs& a, const TablesRT& t, cudaStream_t st) {
    if (a.nelem <= 0) return 0;
    if (dim == 3 && a.geom_diag && !a.zc) {
        // fast path: register-slab kernel (axis-aligned hexes, CF_CONST / CF_K, Q1-Q2)
        if (launch_form_apply_slab(p, op, cf, a, t, st) == 0) return 0;
    }
    if (dim == 3) {
        switch (p) {
            case 1: return launch_form_apply_h1(op, cf, a, t, st);
            case 2: return launch_form_apply_h2(op, cf, a, t, st);
            case 3: return launch_form_apply_h3(op, cf, a, t, st);
            case 4: return launch_form_apply_h4(op, cf, a, t, st);
        }
    } else if (dim == 2) {
        switch (p) {
            case 1: return launch_form_apply_q1(op, cf, a, t, st);
            case 2: return launch_form_apply_q2(op, cf, a, t, st);
            case 3: return launch_form_apply_q3(op, cf, a, t, st);
            case 4: return launch_form_apply_q4(op, cf, a, t, st);
        }
    }
    return 3;
}

template <int OP, int CF>
static int diag_one(const FormArgs& a, const TablesRT& t, cudaStream_t st) {
    constexpr int NF = CoefNF<CF, OP>::value;
    constexpr int NFA = NF > 0 ? NF : 1;
    const int ND = t.D * t.D * t.DZ, NQ = t.Q * t.Q * t.QZ;
    const size_t smem = (size_t)(ND * (1 + NFA) + 11 * NQ) * sizeof(double);
    int grid = a.nelem < 148 * 64 ? a.nelem : 148 * 64;
    form_diag_kernel<OP, CF><<<grid, 128, smem, st>>>(a, t);
    return 0;
}

int launch_form_diag(int dim, int p, int op, int cf, const FormArgs& a, const TablesRT& t, cudaStream_t st) {
    if (a.nelem <= 0) return 0;
    if (dim == 3 && a.geom_diag && launch_form_diag_slab(p, op, cf, a, t, st) == 0) return 0;
#define CASE(OPV, CFV) if (op == OPV && cf == CFV) return diag_one<OPV, CFV>(a, t, st);
    CASE(OP_LIN, CF_CONST) CASE(OP_LIN, CF_LINFIELDS) CASE(OP_LIN, CF_K)
    CASE(OP_JAC, CF_K) CASE(OP_JAC, CF_FULL)
#undef CASE
    return 1;
}

int launch_face(int mode, const FaceArgs& a, cudaStream_t st) {
    if (a.nbdr <= 0) return 0;
    const int grid = (a.nbdr + 127) / 128;
    if (mode == FACE_LOAD) face_kernel<FACE_LOAD><<<grid, 128, 0, st>>>(a);
    else if (mode == FACE_MASS_APPLY) face_kernel<FACE_MASS_APPLY><<<grid, 128, 0, st>>>(a);
    else if (mode == FACE_MASS_DIAG) face_kernel<FACE_MASS_DIAG><<<grid, 128, 0, st>>>(a);
    else return 1;
    return 0;
}

int launch_wall_flux(const WallFluxArgs& a, cudaStream_t st) {
    const int n = a.nfaces * a.nfl;
    if (n <= 0) return 0;
    wall_flux_kernel<<<(n + 127) / 128, 128, 0, st>>>(a);
    return 0;
}

}  // namespace jots
