// Restart files (VisItDataCollection layout of the reference: src/output_manager.cpp:15-21,89-95, src/jots_driver.cpp:214-257).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "mesh.hpp"

namespace jots {

struct RestartData {
    Mesh mesh;
    int order = 0, cycle = 0;
    double time = 0.0;
    std::vector<double> mfem_values;   // Temperature in MFEM's DOF numbering of the H1 space on `mesh`
};

// MFEM's global DOF number of every lattice node of the H1 space (gather: element -> node, lexicographic local order)
void mfem_dof_numbering(const Mesh& m, int p, const std::vector<int>& gather, int64_t ndof, std::vector<int64_t>& mfem_of_node);
void write_mfem_mesh(const std::string& path, const Mesh& m);
// u_nodes: nodal values in the numbering of `space` (= make_space(mesh, p))
void write_restart(const std::string& prefix, int cycle, double time, const Mesh& mesh, const Space& space, const double* u_nodes);
RestartData read_restart(const std::string& prefix, int cycle);
void restart_values_to_nodes(const RestartData& d, const Space& space, std::vector<double>& u_nodes);

}  // namespace jots
