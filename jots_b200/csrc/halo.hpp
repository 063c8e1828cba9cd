// Multi-GPU interface exchange (one subdomain per GPU): sum-exchange of shared-DOF partial sums
// and scalar all-reduce.  Replaces the MPI traffic inside MFEM's P / P^T and hypre's ParCSR
// matvec.  Single-GPU contexts have halo == nullptr.
#pragma once
#include <cstdint>
#include <vector>

namespace jots {
class Ctx;
struct Halo;
struct Comm;
struct Partition;
void comm_unique_id(char out[128]);
Comm* comm_create(const char id[128], int rank, int nranks, int device);
void comm_destroy(Comm* c);
int comm_rank(const Comm* c);
int comm_size(const Comm* c);
Halo* halo_create(const Partition& P, Comm* comm);
void halo_destroy(Halo* h);
void halo_sum_exchange(Ctx& c, double* y);
void halo_allreduce(Ctx& c, double* d_vals, int nv);
void halo_allreduce_max(Ctx& c, double* d_vals, int nv);
int halo_rank(const Ctx& c);
// collective: rank 0 receives (global id, value) of the owned entries of every rank, in rank order (restart output)
void halo_gather_owned(Ctx& c, const double* v, std::vector<int64_t>& gid, std::vector<double>& val);
const int64_t* halo_global_index(Ctx& c);  // device array: local DOF -> global DOF
}  // namespace jots
