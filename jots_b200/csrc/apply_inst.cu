// Instantiations of the generic element kernel for one shape (compile with
// -DJOTS_SHAPE_NAME=h2 -DJOTS_D=3 -DJOTS_Q=4 -DJOTS_DZ=3 -DJOTS_QZ=4).
#include "apply_generic.cuh"
#include "diag_face.cuh"
#include "launch.hpp"

namespace jots {

#define JOTS_CAT2(a, b) a##b
#define JOTS_CAT(a, b) JOTS_CAT2(a, b)

template <int D, int Q, int DZ, int QZ>
static Tables<D, Q, DZ, QZ> make_tables(const TablesRT& r) {
    Tables<D, Q, DZ, QZ> t;
    for (int i = 0; i < Q * D; ++i) { t.B[i] = r.B[i]; t.G[i] = r.G[i]; }
    for (int i = 0; i < Q * Q; ++i) t.Gq[i] = r.Gq[i];
    for (int i = 0; i < Q; ++i) t.W[i] = r.W[i];
    for (int i = 0; i < QZ * DZ; ++i) { t.Bz[i] = r.Bz[i]; t.Gz[i] = r.Gz[i]; }
    for (int i = 0; i < QZ * QZ; ++i) t.Gqz[i] = r.Gqz[i];
    for (int i = 0; i < QZ; ++i) t.Wz[i] = r.Wz[i];
    return t;
}

template <int D, int Q, int DZ, int QZ, int OP, int CF>
static int launch_one(const FormArgs& a, const TablesRT& r, cudaStream_t st) {
    constexpr int NE = DefaultNE<Q>::value;
    const Tables<D, Q, DZ, QZ> t = make_tables<D, Q, DZ, QZ>(r);
    const dim3 block(Q, Q, NE);
    const int grid = (a.nelem + NE - 1) / NE;
    const size_t smem = (size_t)NE * 2 * QZ * Q * Q * sizeof(double);
    form_apply_kernel<D, Q, DZ, QZ, OP, CF, NE><<<grid, block, smem, st>>>(a, t);
    static const std::string label = kernel_label("form_apply_kernel", D, Q, OP, CF, ("," + std::to_string(DZ) + "," + std::to_string(QZ) + "," + std::to_string(NE)).c_str());
    last_element_kernel() = label.c_str();
    return 0;
}

int JOTS_CAT(launch_form_apply_, JOTS_SHAPE_NAME)(int op, int cf, const FormArgs& a, const TablesRT& t, cudaStream_t st) {
    constexpr int D = JOTS_D, Q = JOTS_Q, DZ = JOTS_DZ, QZ = JOTS_QZ;
    if (t.D != D || t.Q != Q || t.DZ != DZ || t.QZ != QZ) return 2;
#define CASE(OPV, CFV) if (op == OPV && cf == CFV) return launch_one<D, Q, DZ, QZ, OPV, CFV>(a, t, st);
    CASE(OP_LIN, CF_CONST) CASE(OP_LIN, CF_LINFIELDS) CASE(OP_LIN, CF_K)
    CASE(OP_JAC, CF_K) CASE(OP_JAC, CF_FULL)
    CASE(OP_RES, CF_CONST) CASE(OP_RES, CF_K) CASE(OP_RES, CF_FULL)
#undef CASE
    return 1;
}

}  // namespace jots
