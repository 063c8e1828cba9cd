// Domain decomposition, one subdomain per GPU (replaces ParMesh(comm, mesh) + ParFiniteElementSpace
// group communication, /root/reference/src/jots_driver.cpp:178,196).  Every rank builds the same
// partition redundantly from the global space (deterministic, no communication):
//   * elements: box partition of structured bricks (2 -> 2x1x1, 4 -> 2x2x1, 8 -> 2x2x2, ...) or
//     recursive coordinate bisection of element centroids;
//   * a subdomain holds every DOF its elements touch ("consistent local vector"); a DOF is owned by
//     the lowest rank touching it (MFEM's group-master rule); local numbering = owned DOFs first
//     (ascending global id), then the shared DOFs owned elsewhere;
//   * per neighbour: the shared DOFs in ascending global id (identical order on both sides).
#pragma once
#include <cstdint>
#include <vector>

#include "mesh.hpp"

namespace jots {

struct Partition {
    int rank = 0, nranks = 1;
    int64_t n_global = 0, n_local = 0, n_owned = 0;
    Space local;                                 // local space (gather in local numbering)
    std::vector<int64_t> global_id;              // [n_local]
    std::vector<int64_t> local_elems;            // global element ids of the local elements
    std::vector<int> nbr_rank;                   // ascending
    std::vector<std::vector<int>> nbr_dofs;      // local indices shared with each neighbour
    // canonical-order reduction of shared DOFs: for shared DOF s (local index red_dof[s]) the
    // contributions red_src[red_ptr[s]..red_ptr[s+1]) are summed left to right; an entry >= 0 indexes
    // the concatenated receive buffer, -1 is the rank's own partial sum (ascending contributor rank)
    std::vector<int> red_dof, red_ptr, red_src;
    std::vector<int> send_idx;                   // concatenated per-neighbour send lists (local indices)
    std::vector<int> nbr_offset;                 // [n_nbr+1] offsets into send/recv buffers
};

// element -> rank; dims (px,py,pz) optional for bricks
void partition_elements(const Mesh& m, int nranks, const int* dims, std::vector<int>& elem_rank);
Partition make_partition(const Space& global, const std::vector<int>& elem_rank, int nranks, int rank);
// the caller's decomposition (partition.cpp); local index of caller L-dof l: position of ldof_global_id[l] in global_id
Partition make_partition_from_tables(int dim, int p, int64_t n_ldof, int64_t n_elem, const int* gather, const double* elem_corner,
                                     int64_t n_bdr, const int* bdr_gather, const double* bdr_corner, const int* bdr_attr,
                                     const int64_t* ldof_global_id, int64_t n_true, const int* ldof_of_true,
                                     int n_nbr, const int* nbr_rank, const int64_t* nbr_ptr, const int* nbr_ldofs,
                                     int n_attributes, const int* attributes, int rank, int nranks);

}  // namespace jots
