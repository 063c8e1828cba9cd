// Multi-GPU interface exchange (one subdomain per GPU): sum-exchange of shared-DOF partial sums
// and scalar all-reduce.  Replaces the MPI traffic inside MFEM's P / P^T and hypre's ParCSR
// matvec.  Single-GPU contexts have halo == nullptr.
#pragma once
#include <cstdint>

namespace jots {
class Ctx;
struct Halo;
void halo_sum_exchange(Ctx& c, double* y);
void halo_allreduce(Ctx& c, double* d_vals, int nv);
const int64_t* halo_global_index(Ctx& c);  // device array: local DOF -> global DOF
}  // namespace jots
