// Host-visible plain structs shared with the loop-based kernels (no device code here).
#pragma once
#include "forms.hpp"

namespace jots {

struct TablesRT {
    int D, Q, DZ, QZ;
    double B[30], G[30], Gq[36], W[6];
    double Bz[30], Gz[30], Gqz[36], Wz[6];
};

enum { FACE_LOAD = 0, FACE_MASS_APPLY = 1, FACE_MASS_DIAG = 2 };

struct FaceArgs {
    int dim, D, nq, nfl, nbdr;
    const int* bdr_gather;      // [nbdr*nfl]
    const double* bdr_corner;   // [nbdr*2^(dim-1)*dim]
    const int* bdr_attr;        // [nbdr]
    const double* g_attr;       // flux per attribute index (attr-1), 0 for unmapped/essential
    const double* const* g_nodal;  // per attribute index: nodal flux field (preCICE heat-flux BC: GridFunctionCoefficient) or null
    int n_attr;
    const unsigned char* ess;   // [ndof] essential marker (may be null)
    const double* z;            // state for rho_h, C_h (null -> uniform: divide by div_const)
    const double* x;            // FACE_MASS_APPLY input
    double* y;
    int mask_in, mask_out;
    double scale;
    double div_const;           // constant divisor rho*C (1 for the plain Neumann vector)
    MatSet mat;
    double Bf[25];              // [nq*D] basis at the face rule
    double Wf[5];
    double xq[5];               // 1-D rule points (for the surface Jacobian)
};

// wall heat flux -k(T) grad T . n / |n| at the nodes of boundary elements (PreciceIsothermalBC::RetrieveWriteData,
// /root/reference/src/boundary_condition.cpp:152-191): affine elements
struct WallFluxArgs {
    int dim, D, nloc, nfl, nfaces;
    const int* faces;          // boundary elements of the BC
    const int* bdr_elem;       // adjacent element of every boundary element
    const int* node_loc;       // [nbdr*nfl] local lattice index of every boundary-element node in that element
    const int* bdr_face;       // its local face (axis = f / 2, side = f % 2)
    const int* gather;         // [nelem][nloc], essential DOFs complemented
    const double* geom;        // invJ + det per element (stride 10) or one record (stride 0)
    int geom_stride;
    const double* u;
    double* out;               // [nfaces][nfl]
    Poly k;
    double Dn[25];             // Dn[i*D + m] = l_m'(x_i): derivative of the nodal basis at the nodes
};

}  // namespace jots
