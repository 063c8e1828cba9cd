// Host-side mesh + H1 lattice space: the tables MFEM's Mesh / ParFiniteElementSpace hand to the
// conduction operators in the reference (/root/reference/src/jots_driver.cpp:143-211):
// element -> DOF gather table, element geometry, boundary faces with attributes, essential DOFs.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace jots {

struct Mesh {
    int dim = 0;
    std::vector<double> verts;      // [nv*dim]
    std::vector<int64_t> elems;     // [ne*2^dim]  MFEM local vertex order
    std::vector<int> elem_attr;     // [ne]
    std::vector<int64_t> bdr;       // [nb*2^(dim-1)]
    std::vector<int> bdr_attr;      // [nb]
    bool is_brick = false;          // structured hint (make_brick)
    int64_t bn[3] = {0, 0, 0};
    double bl[3] = {0, 0, 0};

    int nvpe() const { return 1 << dim; }
    int nvpf() const { return 1 << (dim - 1); }
    int64_t n_elem() const { return (int64_t)elems.size() / nvpe(); }
    int64_t n_bdr() const { return (int64_t)bdr.size() / nvpf(); }
    int64_t n_verts() const { return (int64_t)verts.size() / dim; }
    std::vector<int> bdr_attributes() const;  // sorted unique
};

// corner bits (a,b,c) -> MFEM local vertex id
int corner_vertex(int dim, int a, int b, int c);

Mesh read_mfem_mesh(const std::string& path);   // Mesh(file,1): fix_orientation=true
Mesh make_brick(int64_t nx, int64_t ny, int64_t nz, double lx, double ly, double lz);
Mesh make_rect(int64_t nx, int64_t ny, double lx, double ly, double x0, double y0);
Mesh uniform_refine(const Mesh& m);             // Mesh::UniformRefinement
int64_t fix_orientation(Mesh& m);

// lattice numbering of (p+1)^dim nodes per element shared on vertices/edges/faces, numbered in
// order of first appearance (elements in order, local nodes lexicographic, x fastest)
void lattice_numbering(const Mesh& m, int p, std::vector<int64_t>& gather, int64_t& nnodes);
void brick_lattice_numbering(int64_t nx, int64_t ny, int64_t nz, int p, std::vector<int64_t>& gather, int64_t& nnodes);

void face_lattice_nodes(int dim, int p, int f, std::vector<int>& out);
void face_vertices_local(int dim, int f, std::vector<int>& out);
void locate_boundary_faces(const Mesh& m, std::vector<int64_t>& be, std::vector<int>& bf);

struct Space {
    int dim = 0, p = 0, D = 0, Q = 0, nloc = 0, nfl = 0;
    int64_t nelem = 0, ndof = 0, nbdr = 0;
    std::vector<int> gather;          // [nelem*nloc]
    std::vector<double> elem_corner;  // [nelem*2^dim*dim] vertex coordinates in corner-bit order
    std::vector<int64_t> bdr_elem;
    std::vector<int> bdr_face, bdr_attr;
    std::vector<int> bdr_gather;      // [nbdr*nfl]
    std::vector<double> bdr_corner;   // [nbdr*2^(dim-1)*dim] face corners, lower free axis fastest
    std::vector<int> attributes;      // sorted unique boundary attributes
    std::vector<double> nodes;        // GL nodes on [0,1]
    bool is_brick = false;
    int64_t bn[3] = {0, 0, 0};
    double bl[3] = {0, 0, 0};
    // subdomain spaces: DOFs of the GLOBAL boundary elements of attribute attributes[i] (local ids)
    bool has_attr_dofs = false;
    bool sync_bdr_markers = false;   // caller-supplied subdomain: complete the boundary DOF sets by a marker exchange (Ctx::set_bcs)
    std::vector<std::vector<int>> attr_dofs;

    void node_coordinates(std::vector<double>& out) const;  // [ndof*dim]
    void essential_dofs(const std::vector<int>& attrs, std::vector<int>& out) const;  // sorted unique
};

Space make_space(const Mesh& m, int p);

// Boundary-element adjacency through the DOF ids (spaces built from caller-supplied tables carry none): bdr_elem / bdr_face of
// every boundary element whose adjacent element is unknown, found by matching the face's corner DOFs against the element faces.
void fill_boundary_adjacency(Space& s);
// loc[b * nfl + m] = local lattice index, in the adjacent element, of the DOF bdr_gather[b * nfl + m] (any face DOF order);
// -1 where the adjacency is unknown
void boundary_node_locals(const Space& s, std::vector<int>& loc);

// MFEM's native local DOF order of the H1 tensor elements (H1_QuadrilateralElement / H1_HexahedronElement::GetDofMap,
// fem/fe/fe_h1.cpp of MFEM 4.7): vertices, then edge interiors (edge by edge, each in the direction of
// increasing coordinate on hexes; quads run edges 2 and 3 backwards), then face interiors, then the element interior.
// map[lexicographic index] = native index, i.e. what ParFiniteElementSpace::GetElementDofs returns at position map[l] is the
// DOF of lexicographic node l = i + (p+1) (j + (p+1) k).  Restated from the MFEM sources (not in /root/reference): a binding
// should assert it against GetDofMap() of the element it actually uses.
void h1_dof_map(int dim, int p, std::vector<int>& map);

// Element patches of the unique-pencil slab kernel (apply_patch.cuh): PX x PY x 1 lattice-adjacent hexes with aligned local
// axes, their (D-1)PX+1 x (D-1)PY+1 x D nodes listed once.  Record (ints): nodes [k][J][I] (essential DOFs as ~index,
// 0x80000000 where no element of the patch touches the slot), then PX*PY element ids (-1: hole), padded to a multiple of 4.
// The lattice is recovered from the gather table (any L-vector numbering).  Returns the number of patches, 0 when the mesh
// is not a conforming lattice; *fill = nelem / (patches * PX * PY).
int64_t make_patches(const Space& s, const std::vector<int>& ess, int PX, int PY, std::vector<int>& tab, double* fill);

}  // namespace jots
