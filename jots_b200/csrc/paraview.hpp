// ParaView output in the layout of the reference's ParaViewDataCollection (OutputManager, /root/reference/src/output_manager.cpp:
// 23-28: collection "ParaView", levels of detail = FE order, binary, high-order output; WriteVizOutput :81-87).
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "mesh.hpp"

namespace jots {

struct VizField {
    std::string name;
    const double* nodes;    // nodal values in the numbering of the space
};

// lex_of_vtk[j] = lexicographic index (i + (p+1)(j + (p+1)k)) of node j of VTK_LAGRANGE_QUADRILATERAL / _HEXAHEDRON
// (vtkLagrangeQuadrilateral / vtkLagrangeHexahedron::PointIndexFromIJK, the VTK >= 9 / file version 2.2 edge order)
void vtk_lagrange_order(int dim, int p, std::vector<int>& lex_of_vtk);

// One collection: <dir>/<name>.pvd + <dir>/Cycle<cycle:06d>/{data.pvtu, proc000000.vtu}; dir = "ParaView", name = "ParaView"
// in the reference.  restart_mode (UseRestartMode): the first save keeps the entries of an existing .pvd that lie before the
// current time instead of starting a new collection.
struct ParaViewCollection {
    std::string dir = "ParaView", name = "ParaView";
    bool restart_mode = false;
    bool append = false;       // first save keeps every entry of an existing .pvd (a stateless caller adding one cycle)
    bool opened = false;
    std::vector<std::pair<double, std::string>> entries;   // (time, file relative to dir)

    void save(int cycle, double time, const Space& space, const std::vector<int>& elem_attr, const std::vector<VizField>& fields);
};

}  // namespace jots
