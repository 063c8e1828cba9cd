// Fused BLAS-1 kernels for the Krylov / Newton loops (replace mfem::Vector ops and Dot inside
// CGSolver / GMRESSolver / FGMRESSolver / NewtonSolver, /root/reference/src/jots_common.inl:11-18).
// All are HBM-bound streaming kernels: 128-bit accesses where alignment allows, grid sized to a
// multiple of the SM count, dot products reduced with warp shuffles -> shared -> one partial per
// block, final sum by the last block to finish (deterministic order).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "vec.hpp"

namespace jots {


__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// reduce NV per-thread values over the grid; result[i] written by the last block
template <int NV>
__device__ __forceinline__ void grid_reduce(double (&v)[NV], DotScratch s, double* result) {
    __shared__ double sh[NV][VEC_THREADS / 32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = warp_sum(v[i]);
        if (lane == 0) sh[i][wid] = v[i];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double t = lane < VEC_THREADS / 32 ? sh[i][lane] : 0.0;
            t = warp_sum(t);
            if (lane == 0) s.partial[(size_t)blockIdx.x * NV + i] = t;
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int ticket = atomicAdd(s.counter, 1u);
        last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double t = 0.0;
            for (int b = threadIdx.x; b < (int)gridDim.x; b += VEC_THREADS) t += s.partial[(size_t)b * NV + i];
            t = warp_sum(t);
            __syncthreads();
            if (lane == 0) sh[i][wid] = t;
            __syncthreads();
            if (wid == 0) {
                double u = lane < VEC_THREADS / 32 ? sh[i][lane] : 0.0;
                u = warp_sum(u);
                if (lane == 0) result[i] = u;
            }
        }
        if (threadIdx.x == 0) *s.counter = 0u;
    }
}

}  // namespace jots
