#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "forms.hpp"

namespace jots {
constexpr int VEC_THREADS = 256;
constexpr int DOT_MAX_BLOCKS = 148 * 8;
struct DotScratch {
    double* partial;        // [DOT_MAX_BLOCKS * 4]
    unsigned int* counter;  // zero-initialised ticket
};
void v_waxpby(cudaStream_t st, int64_t n, double a, const double* x, double b, const double* y, double* w);
void v_axpy(cudaStream_t st, int64_t n, double a, const double* x, double* y);
void v_scale(cudaStream_t st, int64_t n, double a, const double* x, double* y);
void v_mul(cudaStream_t st, int64_t n, const double* d, const double* x, double* y);
void v_mul_waxpby(cudaStream_t st, int64_t n, const double* d, double a, const double* x, double b, const double* z, double* y);
void v_fill(cudaStream_t st, int64_t n, double a, double* y);
void v_rsqrt(cudaStream_t st, int64_t n, const double* d, double* y);
void v_recip(cudaStream_t st, int64_t n, const double* d, double* y);
void v_hash_unit(cudaStream_t st, int64_t n, const int64_t* gidx, int64_t offset, double* y);
void v_poly_field(cudaStream_t st, int64_t n, const Poly& P, int order, const double* u, double* out);   // property field / d/dT / d2/dT2
void v_set_idx(cudaStream_t st, int n, const int* idx, double v, double* y);
void v_copy_idx(cudaStream_t st, int n, const int* idx, const double* x, double* y);      // y[idx[i]] = x[idx[i]]
void v_scatter_idx(cudaStream_t st, int n, const int* idx, const double* x, double* y);   // y[idx[i]] = x[i]
void v_pack(cudaStream_t st, int n, const int* idx, const double* y, double* out);
void v_reduce_shared(cudaStream_t st, int n, const int* dof, const int* ptr, const int* src, const double* recv, double* y);
void v_max(cudaStream_t st, int64_t n, const double* x, DotScratch s, double* result);
void v_mgs_step(cudaStream_t st, int64_t n, int64_t n_owned, double* r, const double* Vprev, const double* hprev, const double* Vk,
                DotScratch s, double* result);
void v_mul_axpby(cudaStream_t st, int64_t n, const double* d, double a, const double* x, double b, const double* z, double* y);
void v_mul2(cudaStream_t st, int64_t n, const double* d, const double* x, double c, double* y1, double* y2);
void v_dots(cudaStream_t st, int64_t n, int nv, const double* const* a, const double* const* b, DotScratch s, double* result);
void v_cg_update(cudaStream_t st, int64_t n, int64_t n_owned, double alpha, const double* d, const double* Ad, double* x, double* r,
                 const double* dinv, double* z, DotScratch s, double* result);
// one-block modified Gram-Schmidt sweep for small problems: out[k] = (r, V_k), r -= out[k] V_k (k < nv), out[nv] = (r, r)
void v_mgs_sweep_small(cudaStream_t st, int n, double* r, const double* const* V, int nv, double* out);
// device-resident CG scalars (vec.cu: S layout)
void v_cgd_update(cudaStream_t st, int64_t n, int64_t n_owned, double* S, int i, const double* d, const double* Ad, double* x, double* r,
                  const double* dinv, double* z, DotScratch s);
void v_cgd_direction(cudaStream_t st, int64_t n, double* S, int i, const double* zz, const double* d, double* dout, bool update, bool zero_den);
void v_zero_essdot(cudaStream_t st, int64_t n, double* y, int n_idx, const int* idx, const double* x, double* dot, const double* skip);
void v_cgd_dot(cudaStream_t st, int64_t n_owned, double* S, const double* a, const double* b, DotScratch s);
}  // namespace jots
