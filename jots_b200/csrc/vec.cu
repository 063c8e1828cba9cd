#include <cstdint>
#include <cstdlib>
#include "vec.cuh"
#include "vec.hpp"
#include "coef.cuh"

namespace jots {

static inline int grid_for(int64_t n, int per_thread = 4) {
    int64_t b = (n + (int64_t)VEC_THREADS * per_thread - 1) / ((int64_t)VEC_THREADS * per_thread);
    if (b < 1) b = 1;
    if (b > DOT_MAX_BLOCKS) b = DOT_MAX_BLOCKS;
    return (int)b;
}

// w = a*x + b*y   (x, y may alias w)
__global__ void k_waxpby(int64_t n, double a, const double* __restrict__ x, double b, const double* __restrict__ y, double* w) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) w[i] = a * x[i] + b * y[i];
}
__global__ void k_axpy(int64_t n, double a, const double* __restrict__ x, double* y) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] += a * x[i];
}
__global__ void k_scale(int64_t n, double a, const double* __restrict__ x, double* y) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = a * x[i];
}
__global__ void k_mul(int64_t n, const double* __restrict__ d, const double* __restrict__ x, double* y) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = d[i] * x[i];
}
// y = d * (x + c * r * d')  helpers for Chebyshev:  y = a*d*x + b*d*z
__global__ void k_mul_waxpby(int64_t n, const double* __restrict__ d, double a, const double* __restrict__ x, double b,
                             const double* __restrict__ z, double* y) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = d[i] * (a * x[i] + b * z[i]);
}
__global__ void k_fill(int64_t n, double a, double* y) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = a;
}
__global__ void k_rsqrt(int64_t n, const double* __restrict__ d, double* y) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = 1.0 / sqrt(d[i]);
}
__global__ void k_recip(int64_t n, const double* __restrict__ d, double* y) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = 1.0 / d[i];
}
// deterministic pseudo-random vector in [-1,1): splitmix64 of the global index
__global__ void k_hash_unit(int64_t n, const int64_t* __restrict__ gidx, int64_t offset, double* y) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t gi = (uint64_t)(gidx ? gidx[i] : i) + (uint64_t)offset + 1ull;
        uint64_t z = gi * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z = z ^ (z >> 31);
        y[i] = (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
    }
}
__global__ void k_set_idx(int n, const int* __restrict__ idx, double v, double* y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[idx[i]] = v;
}
__global__ void k_scatter_idx(int n, const int* __restrict__ idx, const double* __restrict__ x, double* y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[idx[i]] = x[i];
}
__global__ void k_copy_idx(int n, const int* __restrict__ idx, const double* __restrict__ x, double* y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[idx[i]] = x[idx[i]];
}

// nodal material-property field (order 0) or its first / second derivative with respect to T at the state u: what
// OutputManager::UpdateGridFunctions projects for the registered property coefficients
// (/root/reference/src/output_manager.cpp:63-73, src/jots_driver.cpp:343, src/material_property.cpp:36-81)
__global__ void k_poly_field(int64_t n, Poly P, int order, const double* __restrict__ u, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double val, d1, d2;
        poly_eval(P, u ? u[i] : 0.0, val, d1, d2);
        out[i] = order == 0 ? val : order == 1 ? d1 : d2;
    }
}

// interface exchange helpers
__global__ void k_pack(int n, const int* __restrict__ idx, const double* __restrict__ y, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = y[idx[i]];
}
// canonical-order sum of the contributions of every sharing rank (ascending rank; -1 = own value)
__global__ void k_reduce_shared(int n, const int* __restrict__ dof, const int* __restrict__ ptr, const int* __restrict__ src,
                                const double* __restrict__ recv, double* y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = dof[i];
    const double own = y[d];
    double acc = 0.0;
    for (int k = ptr[i]; k < ptr[i + 1]; ++k) { const int s = src[k]; acc += s < 0 ? own : recv[s]; }
    y[d] = acc;
}

template <int NV>
__global__ void __launch_bounds__(VEC_THREADS) k_dots(int64_t n, const double* __restrict__ a0, const double* __restrict__ b0,
                                                       const double* __restrict__ a1, const double* __restrict__ b1,
                                                       const double* __restrict__ a2, const double* __restrict__ b2,
                                                       DotScratch s, double* result) {
    double v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        v[0] += a0[i] * b0[i];
        if (NV > 1) v[1] += a1[i] * b1[i];
        if (NV > 2) v[2] += a2[i] * b2[i];
    }
    grid_reduce<NV>(v, s, result);
}

// CG update fused: x += alpha d ; r -= alpha z ; (optional) zz = dinv*r ; result = (r, zz or r)
__global__ void __launch_bounds__(VEC_THREADS) k_cg_update(int64_t n, int64_t n_owned, double alpha, const double* __restrict__ d,
                                                            const double* Ad, double* x, double* r,
                                                            const double* __restrict__ dinv, double* z, DotScratch s, double* result) {
    double v[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        x[i] += alpha * d[i];
        const double ri = r[i] - alpha * Ad[i];
        r[i] = ri;
        double zi = ri;
        if (dinv) { zi = dinv[i] * ri; z[i] = zi; }
        if (i < n_owned) v[0] += ri * zi;
    }
    grid_reduce<1>(v, s, result);
}

// ---- device-resident CG: the scalars of mfem::CGSolver::Mult live in device memory, so an iteration needs no host round trip.
// S[0], S[1] = (B r, r) ping-pong: iteration i reads nom = S[(i+1)&1] and writes betanom = S[i&1];  S[2] = den = (d, A d);
// S[3] = r0 (stopping threshold);  S[4] = stop code (0 running, 1 converged, 2 negative (B r, r), 3 den == 0);
// S[5] = iteration at which it stopped;  S[6] = last (B r, r);  S[7] = alpha of the last update (side-stream loop).  Every decision is a function of fully reduced scalars that
// were final before the kernel started, so all blocks of a kernel take the same branch (grid_reduce needs that).
// SIDE: x += alpha d is not done here but handed to the next operator apply as its side stream (FormArgs::side_*), which
// reads alpha from S[7]; the loop's epilogue applies the one update that is still pending when the iteration stops.
// Measured dead end (profiles/README.md r3j): L2 residency hints between the two vector kernels -- the update kernel streaming r,
// A d, D^-1 as evict-first and leaving z behind as evict-last lines (createpolicy + st.global.L2::cache_hint) so that the direction
// kernel's read of z is served from L2: 318.2 -> 319.8 ms per step.
template <bool SIDE>
__global__ void __launch_bounds__(VEC_THREADS) k_cgd_update(int64_t n, int64_t n_owned, double* S, int i, const double* __restrict__ d,
                                                             const double* Ad, double* x, double* r, const double* __restrict__ dinv,
                                                             double* z, DotScratch s) {
    if (S[4] != 0.0) return;
    const double den = S[2], nom = S[(i + 1) & 1];
    if (den == 0.0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) { S[7] = 0.0; S[5] = (double)i; S[6] = nom; __threadfence(); S[4] = 3.0; }
        return;
    }
    const double alpha = nom / den;
    if (SIDE && blockIdx.x == 0 && threadIdx.x == 0) S[7] = alpha;
    double v[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        if (!SIDE) x[k] += alpha * d[k];
        const double ri = r[k] - alpha * Ad[k];
        r[k] = ri;
        double zi = ri;
        if (dinv) { zi = dinv[k] * ri; z[k] = zi; }
        if (k < n_owned) v[0] += ri * zi;
    }
    grid_reduce<1>(v, s, &S[i & 1]);
}
// convergence test of iteration i, then dout = zz + beta d (update == 0: test only, the last iteration); dout may be d
__global__ void __launch_bounds__(VEC_THREADS) k_cgd_direction(int64_t n, double* S, int i, const double* __restrict__ zz, const double* d, double* dout,
                                                                int update) {
    if (S[4] != 0.0) return;
    const double betanom = S[i & 1], nom = S[(i + 1) & 1];
    const bool first = blockIdx.x == 0 && threadIdx.x == 0;
    if (first && update == 2) S[2] = 0.0;     // the operator apply that follows accumulates (d, A d) here (k_zero_essdot, FormArgs::dot)
    if (betanom < 0.0 || betanom <= S[3]) {
        if (first) { S[5] = (double)i; S[6] = betanom; __threadfence(); S[4] = betanom < 0.0 ? 2.0 : 1.0; }
        return;
    }
    if (first) S[6] = betanom;
    if (!update) return;
    const double beta = betanom / nom;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) dout[k] = zz[k] + beta * d[k];
}
__global__ void __launch_bounds__(VEC_THREADS) k_cgd_dot(int64_t n, double* S, const double* __restrict__ a, const double* __restrict__ b, DotScratch s) {
    if (S[4] != 0.0) return;
    double v[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) v[0] += a[k] * b[k];
    grid_reduce<1>(v, s, &S[2]);
}
void v_cgd_update(cudaStream_t st, int64_t n, int64_t n_owned, double* S, int i, const double* d, const double* Ad, double* x, double* r,
                  const double* dinv, double* z, DotScratch s) {
    if (x) k_cgd_update<false><<<grid_for(n, 8), VEC_THREADS, 0, st>>>(n, n_owned, S, i, d, Ad, x, r, dinv, z, s);
    else k_cgd_update<true><<<grid_for(n, 8), VEC_THREADS, 0, st>>>(n, n_owned, S, i, nullptr, Ad, nullptr, r, dinv, z, s);
}
void v_cgd_direction(cudaStream_t st, int64_t n, double* S, int i, const double* zz, const double* d, double* dout, bool update, bool zero_den) {
    k_cgd_direction<<<update ? grid_for(n) : 1, VEC_THREADS, 0, st>>>(n, S, i, zz, d, dout, update ? (zero_den ? 2 : 1) : 0);
}
// y = 0 and *dot += sum over the listed (owned, essential) entries of x^2: what the unit-diagonal rows of an eliminated operator
// contribute to (x, A x).  Precedes an element kernel that accumulates the rest (FormArgs::dot).  skip as in FormArgs.
__global__ void __launch_bounds__(VEC_THREADS) k_zero_essdot(int64_t n, double* __restrict__ y, int n_idx, const int* __restrict__ idx,
                                                              const double* __restrict__ x, double* dot, const double* skip) {
    if (skip && *skip != 0.0) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t head = (reinterpret_cast<uintptr_t>(y) & 8) ? 1 : 0;     // work vectors are only 8-byte aligned
    double2* y2 = reinterpret_cast<double2*>(y + head);
    const int64_t m2 = (n - head) / 2;
    for (int64_t k = t0; k < m2; k += stride) y2[k] = make_double2(0.0, 0.0);
    if (t0 == 0 && n > 0) { if (head) y[0] = 0.0; if ((n - head) & 1) y[n - 1] = 0.0; }
    double v = 0.0;
    for (int64_t k = t0; k < n_idx; k += stride) { const double xv = x[idx[k]]; v += xv * xv; }
    __shared__ double sh[VEC_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < VEC_THREADS / 32; ++w) v += sh[w];
        if (v != 0.0) atomicAdd(dot, v);
    }
}
void v_zero_essdot(cudaStream_t st, int64_t n, double* y, int n_idx, const int* idx, const double* x, double* dot, const double* skip) {
    // n == 0: only the listed entries (y was zeroed elsewhere)
    k_zero_essdot<<<grid_for(n > 0 ? n : (int64_t)n_idx * 8, 8), VEC_THREADS, 0, st>>>(n, y, n_idx, idx, x, dot, skip);
}
void v_cgd_dot(cudaStream_t st, int64_t n_owned, double* S, const double* a, const double* b, DotScratch s) {
    k_cgd_dot<<<grid_for(n_owned, 8), VEC_THREADS, 0, st>>>(n_owned, S, a, b, s);
}

// max over a vector (JOTSDriver::PostprocessIteration blow-up check, jots_driver.cpp:739): per-block
// maxima, final maximum by the last block to finish
__global__ void __launch_bounds__(VEC_THREADS) k_max(int64_t n, const double* __restrict__ x, DotScratch s, double* result) {
    __shared__ double sh[VEC_THREADS / 32];
    __shared__ bool last;
    double v = -1.7976931348623157e308;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) v = fmax(v, x[i]);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < VEC_THREADS / 32; ++w) v = fmax(v, sh[w]);
        s.partial[blockIdx.x] = v;
        __threadfence();
        last = (atomicAdd(s.counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double m = s.partial[0];
        for (int b = 1; b < (int)gridDim.x; ++b) m = fmax(m, s.partial[b]);
        *result = m;
        *s.counter = 0u;
    }
}
void v_max(cudaStream_t st, int64_t n, const double* x, DotScratch s, double* result) {
    k_max<<<grid_for(n, 8), VEC_THREADS, 0, st>>>(n, x, s, result);
}

// one modified Gram-Schmidt step without a host round trip:
//   r -= (*hprev) * Vprev   (skipped when Vprev == nullptr)
//   *result = (r, Vk) over the owned entries, or (r, r) when Vk == nullptr
// hprev is the device scalar the previous step produced (already all-reduced on multi-GPU runs).
__global__ void __launch_bounds__(VEC_THREADS) k_mgs_step(int64_t n, int64_t n_owned, double* r, const double* __restrict__ Vprev,
                                                           const double* __restrict__ hprev, const double* __restrict__ Vk,
                                                           DotScratch s, double* result) {
    double v[1] = {0.0};
    const double h = Vprev ? *hprev : 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double ri = r[i];
        if (Vprev) { ri -= h * Vprev[i]; r[i] = ri; }
        if (i < n_owned) v[0] += ri * (Vk ? Vk[i] : ri);
    }
    grid_reduce<1>(v, s, result);
}
void v_mgs_step(cudaStream_t st, int64_t n, int64_t n_owned, double* r, const double* Vprev, const double* hprev, const double* Vk,
                DotScratch s, double* result) {
    k_mgs_step<<<grid_for(n, 8), VEC_THREADS, 0, st>>>(n, n_owned, r, Vprev, hprev, Vk, s, result);
}
// The whole modified Gram-Schmidt sweep of one (F)GMRES iteration in ONE block, for problems so small that a launch per basis
// vector is pure latency (the shipped 1 791-DOF cases spent 60 % of their launches here):
//   for k = 0 .. nv-1:  out[k] = (r, V_k);  r -= out[k] V_k;      then out[nv] = (r, r)
// Same operation order as the chained k_mgs_step launches (deterministic block reduction).
__global__ void __launch_bounds__(1024) k_mgs_sweep_small(int n, double* r, const double* const* V, int nv, double* out) {
    __shared__ double sh[32];
    __shared__ double hk;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k = 0; k <= nv; ++k) {
        const double* v = k < nv ? V[k] : r;
        double s = 0.0;
        for (int j = threadIdx.x; j < n; j += blockDim.x) s += r[j] * v[j];
        s = warp_sum(s);
        if (lane == 0) sh[wid] = s;
        __syncthreads();
        if (wid == 0) {
            double t = lane < nw ? sh[lane] : 0.0;
            t = warp_sum(t);
            if (lane == 0) { hk = t; out[k] = t; }
        }
        __syncthreads();
        if (k < nv) {
            const double h = hk;
            for (int j = threadIdx.x; j < n; j += blockDim.x) r[j] -= h * v[j];
        }
        __syncthreads();
    }
}
void v_mgs_sweep_small(cudaStream_t st, int n, double* r, const double* const* V, int nv, double* out) {
    k_mgs_sweep_small<<<1, 1024, 0, st>>>(n, r, V, nv, out);
}

// y = d * x * a + b * z   (Chebyshev Horner step: u = ds * (A t) + c_i * rs)
__global__ void k_mul_axpby(int64_t n, const double* __restrict__ d, double a, const double* __restrict__ x, double b,
                            const double* __restrict__ z, double* y) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = a * d[i] * x[i] + b * z[i];
}
void v_mul_axpby(cudaStream_t st, int64_t n, const double* d, double a, const double* x, double b, const double* z, double* y) {
    k_mul_axpby<<<grid_for(n), VEC_THREADS, 0, st>>>(n, d, a, x, b, z, y);
}

// y1 = d * x ; y2 = c * y1   (first step of the Chebyshev Horner recurrence)
__global__ void k_mul2(int64_t n, const double* __restrict__ d, const double* __restrict__ x, double c, double* y1, double* y2) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = d[i] * x[i];
        y1[i] = v; y2[i] = c * v;
    }
}
void v_mul2(cudaStream_t st, int64_t n, const double* d, const double* x, double c, double* y1, double* y2) {
    k_mul2<<<grid_for(n), VEC_THREADS, 0, st>>>(n, d, x, c, y1, y2);
}

void v_waxpby(cudaStream_t st, int64_t n, double a, const double* x, double b, const double* y, double* w) {
    k_waxpby<<<grid_for(n), VEC_THREADS, 0, st>>>(n, a, x, b, y, w);
}
void v_axpy(cudaStream_t st, int64_t n, double a, const double* x, double* y) { k_axpy<<<grid_for(n), VEC_THREADS, 0, st>>>(n, a, x, y); }
void v_scale(cudaStream_t st, int64_t n, double a, const double* x, double* y) { k_scale<<<grid_for(n), VEC_THREADS, 0, st>>>(n, a, x, y); }
void v_mul(cudaStream_t st, int64_t n, const double* d, const double* x, double* y) { k_mul<<<grid_for(n), VEC_THREADS, 0, st>>>(n, d, x, y); }
void v_mul_waxpby(cudaStream_t st, int64_t n, const double* d, double a, const double* x, double b, const double* z, double* y) {
    k_mul_waxpby<<<grid_for(n), VEC_THREADS, 0, st>>>(n, d, a, x, b, z, y);
}
void v_fill(cudaStream_t st, int64_t n, double a, double* y) { k_fill<<<grid_for(n), VEC_THREADS, 0, st>>>(n, a, y); }
void v_rsqrt(cudaStream_t st, int64_t n, const double* d, double* y) { k_rsqrt<<<grid_for(n), VEC_THREADS, 0, st>>>(n, d, y); }
void v_recip(cudaStream_t st, int64_t n, const double* d, double* y) { k_recip<<<grid_for(n), VEC_THREADS, 0, st>>>(n, d, y); }
void v_hash_unit(cudaStream_t st, int64_t n, const int64_t* gidx, int64_t offset, double* y) {
    k_hash_unit<<<grid_for(n), VEC_THREADS, 0, st>>>(n, gidx, offset, y);
}
void v_set_idx(cudaStream_t st, int n, const int* idx, double v, double* y) {
    if (n > 0) k_set_idx<<<(n + 255) / 256, 256, 0, st>>>(n, idx, v, y);
}
void v_scatter_idx(cudaStream_t st, int n, const int* idx, const double* x, double* y) {
    if (n > 0) k_scatter_idx<<<(n + 255) / 256, 256, 0, st>>>(n, idx, x, y);
}
void v_copy_idx(cudaStream_t st, int n, const int* idx, const double* x, double* y) {
    if (n > 0) k_copy_idx<<<(n + 255) / 256, 256, 0, st>>>(n, idx, x, y);
}
void v_poly_field(cudaStream_t st, int64_t n, const Poly& P, int order, const double* u, double* out) {
    if (n > 0) k_poly_field<<<grid_for(n, 8), VEC_THREADS, 0, st>>>(n, P, order, u, out);
}
void v_pack(cudaStream_t st, int n, const int* idx, const double* y, double* out) {
    if (n > 0) k_pack<<<(n + 255) / 256, 256, 0, st>>>(n, idx, y, out);
}
void v_reduce_shared(cudaStream_t st, int n, const int* dof, const int* ptr, const int* src, const double* recv, double* y) {
    if (n > 0) k_reduce_shared<<<(n + 255) / 256, 256, 0, st>>>(n, dof, ptr, src, recv, y);
}
void v_dots(cudaStream_t st, int64_t n, int nv, const double* const* a, const double* const* b, DotScratch s, double* result) {
    const int g = grid_for(n, 8);
    if (nv == 1) k_dots<1><<<g, VEC_THREADS, 0, st>>>(n, a[0], b[0], nullptr, nullptr, nullptr, nullptr, s, result);
    else if (nv == 2) k_dots<2><<<g, VEC_THREADS, 0, st>>>(n, a[0], b[0], a[1], b[1], nullptr, nullptr, s, result);
    else k_dots<3><<<g, VEC_THREADS, 0, st>>>(n, a[0], b[0], a[1], b[1], a[2], b[2], s, result);
}
void v_cg_update(cudaStream_t st, int64_t n, int64_t n_owned, double alpha, const double* d, const double* Ad, double* x, double* r,
                 const double* dinv, double* z, DotScratch s, double* result) {
    k_cg_update<<<grid_for(n, 8), VEC_THREADS, 0, st>>>(n, n_owned, alpha, d, Ad, x, r, dinv, z, s, result);
}

}  // namespace jots
