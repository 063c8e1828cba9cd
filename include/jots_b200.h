/*
 * jots_b200 -- C ABI of the B200-native implicit conduction solve for JOTS.
 *
 * The reference (j-signorelli/jots) has no C ABI for this path: its conduction operators sit
 * behind C++ virtuals supplied by MFEM.  Each entry point below names the reference interface
 * it replaces (file:line under /root/reference).  INTEGRATION.md shows the binding a JOTS
 * maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success and a non-zero status on failure, or a handle that is
 *     NULL on failure; jots_last_error() returns the message (thread-local).  No exceptions
 *     cross the boundary.  The reference aborts the process on these errors
 *     (MFEM_ABORT / MFEM_VERIFY: src/nl_conduction_operator.cpp:312, src/jots_driver.cpp:741).
 *   - vectors are plain FP64 arrays of jots_ctx_size() entries (true-DOF vectors, the layout of
 *     mfem::Vector u in src/jots_driver.hpp:48).  `loc` says where a pointer lives:
 *     JOTS_HOST (caller-owned host memory: copied to the device and back inside the call) or
 *     JOTS_DEVICE (device pointer on the context's GPU: used in place on the context's stream).
 *   - one context per GPU / rank; there is no CPU fallback: creating a context without a CUDA
 *     device fails.
 */
#ifndef JOTS_B200_H
#define JOTS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jots_mesh_s* jots_mesh_t;     /* host mesh (mfem::Mesh role) */
typedef struct jots_space_s* jots_space_t;   /* host H1 space tables (mfem::ParFiniteElementSpace role) */
typedef struct jots_ctx_s* jots_ctx_t;       /* device context of one subdomain */
typedef struct jots_op_s* jots_op_t;         /* one conduction operator (JOTSIterator) */
typedef struct jots_driver_s* jots_driver_t; /* hot-path callers of JOTSDriver */
typedef struct jots_part_s* jots_part_t;     /* one subdomain of a partitioned space (ParMesh role) */
typedef struct jots_comm_s* jots_comm_t;     /* NCCL communicator, one rank per GPU */

enum { JOTS_HOST = 0, JOTS_DEVICE = 1 };
enum { JOTS_UNIFORM = 0, JOTS_POLYNOMIAL = 1 };                      /* src/option_structure.hpp:25-37 */
enum { JOTS_BC_ISOTHERMAL = 0, JOTS_BC_HEATFLUX = 1, JOTS_BC_SINUSOIDAL_ISOTHERMAL = 2,
       JOTS_BC_SINUSOIDAL_HEATFLUX = 3, JOTS_BC_PRECICE_ISOTHERMAL = 4,
       JOTS_BC_PRECICE_HEATFLUX = 5 };                               /* src/option_structure.hpp:56-71 */
enum { JOTS_CG = 0, JOTS_GMRES = 1, JOTS_FGMRES = 2 };               /* src/option_structure.hpp:91-100 */
enum { JOTS_JACOBI = 0, JOTS_CHEBYSHEV = 1 };                        /* src/option_structure.hpp:102-109 */
enum { JOTS_EULER_IMPLICIT = 0, JOTS_EULER_EXPLICIT = 1, JOTS_RK4 = 2 }; /* src/option_structure.hpp:73-82 */
enum { JOTS_LINEARIZED_UNSTEADY = 0, JOTS_NONLINEAR_UNSTEADY = 1, JOTS_STEADY = 2 }; /* :8-16 */
enum { JOTS_DENSITY = 0, JOTS_SPECIFIC_HEAT = 1, JOTS_THERMAL_CONDUCTIVITY = 2 };    /* :18-23 */
/* building-block forms exposed for parity tests */
enum { JOTS_FORM_MASS = 0,        /* M: linearized int rho C phi phi | nonlinear unit mass, eliminated */
       JOTS_FORM_STIFFNESS = 1,   /* K: int k grad phi . grad phi, not eliminated
                                     (src/linear_conduction_operator.cpp:63,94-107) */
       JOTS_FORM_SYSTEM = 2,      /* A = M + dt K eliminated (:109-120) | Jacobian of R at the state
                                     (src/nl_conduction_operator.cpp:81-90) | steady Jacobian */
       JOTS_FORM_RESIDUAL = 3 };  /* R(k) (src/nl_conduction_operator.cpp:70-79) | steady k(u) */

const char* jots_last_error(void);
int jots_device_count(void);
const char* jots_version(void);

/* ---- mesh: what the driver gets from mfem::Mesh (src/jots_driver.cpp:156-185) ---------------- */
jots_mesh_t jots_mesh_read(const char* path);                       /* Mesh(file, 1) */
jots_mesh_t jots_mesh_brick(int64_t nx, int64_t ny, int64_t nz, double lx, double ly, double lz);
jots_mesh_t jots_mesh_rect(int64_t nx, int64_t ny, double lx, double ly, double x0, double y0);
jots_mesh_t jots_mesh_refine(jots_mesh_t m);                        /* UniformRefinement */
void jots_mesh_destroy(jots_mesh_t m);
int jots_mesh_info(jots_mesh_t m, int* dim, int64_t* n_elem, int64_t* n_bdr, int64_t* n_verts);
int jots_mesh_copy(jots_mesh_t m, double* verts, int64_t* elems, int64_t* bdr, int32_t* bdr_attr);
int jots_mesh_set_vertices(jots_mesh_t m, const double* verts);     /* Mesh::SetVertices: new coordinates [n_verts*dim], same topology */

/* ---- space: H1_FECollection + ParFiniteElementSpace (src/jots_driver.cpp:194-196) ------------- */
jots_space_t jots_space_create(jots_mesh_t m, int order);
/* the same space from tables the caller already holds (ParFiniteElementSpace::GetElementDofs /
   GetBdrElementDofs + vertex coordinates): gather[n_elem][(p+1)^dim] in lexicographic local order (x
   fastest), element corners [n_elem][2^dim][dim] in corner-bit order, boundary faces with their
   (p+1)^(dim-1) DOFs (lower free axis fastest), 2^(dim-1) corners and attribute */
jots_space_t jots_space_create_from_tables(int dim, int order, int64_t n_elem, int64_t n_dof, const int32_t* gather,
                                           const double* elem_corners, int64_t n_bdr, const int32_t* bdr_gather,
                                           const double* bdr_corners, const int32_t* bdr_attr);
/* Host only.  For every boundary element: the element it belongs to, that element's local face (0..2 dim - 1) and, for each of its
   (p+1)^(dim-1) nodes in the CALLER's order, the node's place in the element's (p+1)^dim lattice -- what the wall heat flux of
   PreciceIsothermalBC::RetrieveWriteData (src/boundary_condition.cpp:160-206) needs.  A space built from tables carries no
   adjacency: it is recovered by matching DOF ids.  Any output may be NULL. */
int jots_space_boundary_adjacency(jots_space_t s, int32_t* bdr_elem, int32_t* bdr_face, int32_t* node_loc);
int jots_space_copy_corners(jots_space_t s, double* elem_corners, double* bdr_corners);
void jots_space_destroy(jots_space_t s);
int jots_space_info(jots_space_t s, int* dim, int* order, int64_t* n_elem, int64_t* n_dof, int64_t* n_bdr,
                    int* dofs_per_elem, int* dofs_per_face);
int jots_space_copy_gather(jots_space_t s, int32_t* gather);        /* [n_elem*dofs_per_elem] */
int jots_space_copy_bdr(jots_space_t s, int32_t* bdr_gather, int32_t* bdr_attr, int64_t* bdr_elem, int32_t* bdr_face);
int jots_space_copy_nodes(jots_space_t s, double* xyz);             /* [n_dof*dim] */
/* GetEssentialTrueDofs (src/jots_iterator.cpp:30): returns the count, fills out[] when non-NULL */
int64_t jots_space_essential_dofs(jots_space_t s, const int32_t* attrs, int n_attrs, int32_t* out);
/* Element patches of the unique-pencil element kernel (apply_patch.cuh), exposed so that the index contract can be
   checked bit-exactly on the host: px x py x 1 lattice-adjacent hexes with aligned local axes whose nodes are listed
   once.  Record (int32): ((p)px+1)((p)py+1)(p+1) nodes in [k][J][I] order (essential DOFs as ~index, 0x80000000 where no
   element of the patch touches the slot), then px*py element ids (-1: hole), padded to a multiple of 4.  The lattice is
   recovered from the gather table, whatever the DOF numbering (replaces the per-element gather of
   ParNonlinearForm::Mult / GetGradient, src/nl_conduction_operator.cpp:27-67).  Returns the number of patches (0: not a
   conforming lattice or order < 2), -1 on error; *fill = elements / (patches*px*py); fills out[] when non-NULL. */
int64_t jots_space_patches(jots_space_t s, const int32_t* ess, int64_t n_ess, int px, int py, int* record_ints,
                           double* fill, int32_t* out);

/* Patch size the kernel uses for an order (0 x 0: no patch kernel) and, when items != NULL (room for 128), its phase-3 work
   items: pencil | n_visits << 8 | (element << 5 | local ij) << 10 | (element << 5 | local ij) << 20.  Returns the item count. */
int jots_patch_shape(int order, int* px, int* py, uint32_t* items);

/* MFEM-native -> lexicographic local DOF order (H1_FECollection(order, dim), src/jots_driver.cpp:194-196).
   jots_h1_dof_map: map[l] = position of lexicographic node l = i + (p+1)(j + (p+1)k) in the native order (vertices, edge
   interiors, face interiors, interior) = H1_HexahedronElement / H1_QuadrilateralElement::GetDofMap().  Restated from the MFEM
   4.7 sources, which are not part of the reference tree: a binding should assert it once against GetDofMap().
   jots_gather_from_native: an element-DOF table as FiniteElementSpace::GetElementDofs returns it ([n_elem][(p+1)^dim], native
   order, -1-d for d where MFEM flags a sign) -> the lexicographic table jots_space_create_from_tables and
   jots_partition_create_from_tables expect. */
int jots_h1_dof_map(int dim, int order, int32_t* map);
int jots_gather_from_native(int dim, int order, int64_t n_elem, const int32_t* native, int32_t* lexicographic);

/* ---- device context ----------------------------------------------------------------------------- */
jots_ctx_t jots_ctx_create(jots_space_t s, int device);
void jots_ctx_destroy(jots_ctx_t c);
int64_t jots_ctx_size(jots_ctx_t c);
int jots_ctx_set_stream(jots_ctx_t c, void* cuda_stream);           /* run on the caller's stream */
int jots_ctx_sync(jots_ctx_t c);
/* MaterialProperty objects (src/material_property.cpp:6-81): model + coefficients, highest power first */
int jots_ctx_set_material(jots_ctx_t c, int which, int model, const double* coeffs, int n_coeffs);
/* BoundaryCondition objects, entry i <-> boundary attribute i+1 (src/jots_iterator.cpp:12-25):
   params[4*i..] = value | amplitude, angular frequency, phase, vertical shift */
int jots_ctx_set_bcs(jots_ctx_t c, int n_bc, const int32_t* kinds, const double* params);   /* kinds: JOTS_BC_* */
int64_t jots_ctx_essential_dofs(jots_ctx_t c, int32_t* out);        /* ess_tdof_list */
int jots_ctx_update_bcs(jots_ctx_t c, double time);                 /* BoundaryCondition::UpdateCoeff */
int jots_ctx_apply_dirichlet(jots_ctx_t c, double* u, int loc, int only_time_dependent); /* ProjectBdrCoefficient, src/jots_driver.cpp:684-707 */
int jots_ctx_set_material_state(jots_ctx_t c, const double* u, int loc); /* UpdateAllCoeffs(u), src/jots_driver.cpp:664 */
/* preCICE boundary data path (PreciceBC / PreciceIsothermalBC / PreciceHeatFluxBC, src/boundary_condition.cpp:77-206; the
   adapter that moves the arrays, src/jots_precice.cpp:43-67, stays with the reference).  BC kinds JOTS_BC_PRECICE_ISOTHERMAL /
   JOTS_BC_PRECICE_HEATFLUX (params[0] = default value) keep a nodal coefficient field instead of one value.
     _size / _nodes   the nodes of every boundary element of the BC's attribute, boundary element by boundary element with
                      duplicates kept (bdr_dof_indices, coords: what the adapter registers as the coupling-mesh vertices);
     _set_read_data   coeff_gf.SetSubVector(bdr_dof_indices, read_data_arr) (PreciceBC::UpdateCoeff): Dirichlet values of the
                      isothermal kind, the flux field g of int g_h phi_i dS of the heat-flux kind (already negated by the adapter);
     _write_data      RetrieveWriteData: wall heat flux -k(T) grad T . n / |n| at every listed node from the ADJACENT element's
                      gradient (device kernel, affine elements) for the isothermal kind, the boundary temperatures otherwise. */
int64_t jots_ctx_precice_bc_size(jots_ctx_t c, int bc_index);   /* -1: not a preCICE BC */
int jots_ctx_precice_bc_nodes(jots_ctx_t c, int bc_index, double* coords, int32_t* dofs);
int jots_ctx_precice_bc_set_read_data(jots_ctx_t c, int bc_index, const double* values);
int jots_ctx_precice_bc_write_data(jots_ctx_t c, int bc_index, const double* u, int loc, double* out);
/* Output-side property fields (OutputManager::UpdateGridFunctions -> ProjectCoefficient of the registered
   MaterialProperty coefficient, src/output_manager.cpp:63-73, src/jots_driver.cpp:343): nodal values of property
   `which` (0 rho, 1 C, 2 k), or of its first / second derivative (order 1 / 2), at the material state last set. */
int jots_ctx_material_field(jots_ctx_t c, int which, int order, double* out, int loc);
/* counters: applies, diag assemblies, krylov iters, newton iters, residual evals, steps,
   kernel launches, dots, last linear iters, last linear converged, last newton iters, last newton converged */
int jots_ctx_stats(jots_ctx_t c, int64_t out[12], double norms[2]);
/* element patches of the unique-pencil kernel this context runs (0 patches: the gather-table kernels are in use):
   number of patches, fill = elements / (patches x elements per patch), operator applies that ran the patch kernel */
int jots_ctx_patch_info(jots_ctx_t c, int64_t* n_patches, double* fill, int64_t* applies);
/* operator applies of the conjugate-gradient loop that formed (d, A d) inside the element kernel (no separate pass) */
int jots_ctx_fused_dots(jots_ctx_t c, int64_t* n);
/* ... and of those, the applies that also carried the loop's side stream (x += alpha d of the previous iteration and the zeroing
   of the next output buffer, done by the FP64-bound element kernel while HBM is idle: DESIGN.md 6) */
int jots_ctx_side_applies(jots_ctx_t c, int64_t* n);
/* label of the element kernel the last operator apply launched, e.g. "form_apply_patch_kernel<3,4,OP_JAC,CF_KL,4,4,128>" */
const char* jots_ctx_last_kernel(jots_ctx_t c);
int jots_ctx_reset_stats(jots_ctx_t c);

/* ---- operators (JOTSIterator + TimeDependentOperator) ------------------------------------------- */
/* LinearConductionOperator / NonlinearConductionOperator / SteadyConductionOperator constructors
   (src/linear_conduction_operator.cpp:15-74, src/nl_conduction_operator.cpp:166-278,
   src/steady_conduction_operator.cpp:3-32); solver settings as in Config
   (src/config_file.cpp:131-165).  The material state must be set before for T-dependent properties. */
jots_op_t jots_op_create(jots_ctx_t c, int sim_type, double dt, int time_scheme,
                         int solver, int prec, double rel_tol, double abs_tol, int max_iter,
                         double newton_rel_tol, double newton_abs_tol, int newton_max_iter);
void jots_op_destroy(jots_op_t op);
/* Factory::CreatePrintLevel (src/jots_common.inl:57-91; keys LinearSolverSettings.Print_Level / NewtonSolverSettings.Print_Level,
   src/config_file.cpp:140-141,161-162): "Errors, Warnings, Iterations" -> mask (errors 1, warnings 2, iterations 4, summary 8,
   first-and-last 16; "None" clears, "All" sets all; -1 for an unknown label).  The driver shim applies the config's lists
   itself; operators created with jots_op_create print nothing until a mask is set. */
int jots_print_level_mask(const char* comma_separated_labels);
int jots_op_set_print_level(jots_op_t op, int linear_solver_mask, int newton_mask);
int jots_op_iterate(jots_op_t op, double* u, int loc, double* time_inout);     /* JOTSIterator::Iterate */
int jots_op_implicit_solve(jots_op_t op, double dt, const double* u, double* k, int loc); /* ImplicitSolve */
int jots_op_mult(jots_op_t op, const double* u, double* k, int loc);           /* explicit Mult */
int jots_op_process_mat_prop_update(jots_op_t op, int which);                  /* ProcessMatPropUpdate */
int jots_op_update_neumann(jots_op_t op);                                      /* UpdateNeumann */
int jots_op_get_neumann(jots_op_t op, double* b, int loc);                     /* b_vec */
/* building blocks: y = form(x) at frozen state (state may be NULL where the form has none) */
int jots_op_form_apply(jots_op_t op, int form, const double* x, const double* state, double* y, int loc);
int jots_op_form_diagonal(jots_op_t op, int form, const double* state, double* d, int loc);
/* y = form(x) and *dot = (x, form(x)) over the owned rows, all-reduced: what one conjugate-gradient iteration needs
   (mfem::CGSolver::Mult: den = Dot(d, z) after oper->Mult(d, z)).  *fused = 1 when the element kernel formed the inner product
   in the same pass (two-phase patch kernel), 0 when a separate pass did.  Bilinear forms only (MASS, STIFFNESS, SYSTEM). */
int jots_op_form_apply_dot(jots_op_t op, int form, const double* x, const double* state, double* y, int loc, double* dot, int* fused);
/* The apply one iteration of the device-resident conjugate-gradient loop launches when the element kernel can carry the loop's
   side stream (mfem::CGSolver::Mult: oper->Mult(d, z); den = Dot(d, z); add(x, alpha, d, x) of the PREVIOUS iteration):
   y = form(x), *dot = (x, form(x)), side_x += alpha * side_d, side_zero = 0, one element-kernel launch.  Device pointers
   only (loc = JOTS_DEVICE), 16-byte aligned.  *carried = 0 (and nothing computed) when this form's kernel has no side stream. */
int jots_op_form_apply_side(jots_op_t op, int form, const double* x, const double* state, double* y, int loc, double* dot,
                            double alpha, const double* side_d, double* side_x, double* side_zero, int* carried);
/* repeat y = form(x) `reps` times on device-resident vectors; returns the mean milliseconds per
   apply measured with CUDA events on the context's stream (bench.py's roofline leg) */
int jots_op_form_time(jots_op_t op, int form, const double* x_dev, const double* state_dev, double* y_dev,
                      int reps, double* ms_per_apply);

/* ---- multi-GPU: one subdomain per GPU (ParMesh(comm, mesh), src/jots_driver.cpp:178) ------------- */
/* deterministic partition built redundantly on every rank; dims = {px,py,pz} or NULL (automatic box
   partition of bricks, recursive coordinate bisection otherwise) */
jots_part_t jots_partition_create(jots_mesh_t mesh, jots_space_t global_space, int nranks, int rank, const int32_t* dims);
/* The CALLER's decomposition: what ParMesh(comm, mesh) and ParFiniteElementSpace hold on this rank
   (src/jots_driver.cpp:178,196), so that the binding runs under the reference's own `mpirun -np N` partition.
     gather / bdr_gather      element and boundary-face DOF tables in the caller's L-vector numbering (local lexicographic node
                              order, like jots_space_create_from_tables), bdr_* = the rank's own boundary elements;
     ldof_global_id[n_ldof]   global true-DOF number of every L-vector DOF (ParFiniteElementSpace::GetGlobalTDofNumber);
     ldof_of_true[n_true]     L-vector index of each OWNED true DOF, in the caller's true-DOF order;
     nbr_rank / nbr_ptr / nbr_ldofs   for every rank sharing DOFs with this one: the shared L-vector DOFs (any order, the
                              library sorts both sides by global id);
     bdr_attributes           the GLOBAL boundary attributes (ParMesh::bdr_attributes): a rank need not own a face of each.
   The subdomain is numbered owned DOFs first IN THE CALLER'S TRUE-DOF ORDER (a true-DOF vector is the leading part of a local
   vector: jots_ctx_expand_true), ghosts behind them; jots_partition_copy returns the global ids in that order.  Boundary DOF
   sets (ess_tdof_list, Dirichlet projection) are completed across ranks by a marker exchange in jots_ctx_set_bcs, as
   GetEssentialTrueDofs does. */
jots_part_t jots_partition_create_from_tables(int dim, int order, int64_t n_ldof, int64_t n_elem, const int32_t* gather, const double* elem_corners,
                                              int64_t n_bdr, const int32_t* bdr_gather, const double* bdr_corners, const int32_t* bdr_attr,
                                              const int64_t* ldof_global_id, int64_t n_true, const int32_t* ldof_of_true,
                                              int n_nbr, const int32_t* nbr_rank, const int64_t* nbr_ptr, const int32_t* nbr_ldofs,
                                              int n_bdr_attributes, const int32_t* bdr_attributes, int rank, int nranks);
void jots_partition_destroy(jots_part_t p);
int jots_partition_info(jots_part_t p, int64_t* n_local, int64_t* n_owned, int64_t* n_global, int64_t* n_elem,
                        int* n_neighbors, int64_t* n_shared);
jots_space_t jots_partition_space(jots_part_t p);                     /* local space, owned DOFs first */
int jots_partition_copy(jots_part_t p, int64_t* global_id, int64_t* local_elems);
int64_t jots_partition_neighbor(jots_part_t p, int i, int* rank, int32_t* local_idx); /* shared DOFs, ascending global id */
int64_t jots_partition_reduction(jots_part_t p, int32_t* red_dof, int32_t* red_ptr, int32_t* red_src);
int jots_comm_unique_id(char out[128]);                               /* rank 0, then broadcast by the caller */
jots_comm_t jots_comm_create(const char id[128], int rank, int nranks, int device);
void jots_comm_destroy(jots_comm_t c);
jots_ctx_t jots_ctx_create_part(jots_part_t p, int device, jots_comm_t comm);
/* true-DOF vector (the n_owned leading entries, what mfem::Vector u holds on this rank) -> consistent local vector of
   jots_ctx_size entries: the ghost entries are fetched from their owners (P of ParFiniteElementSpace).  Collective.
   tvec and lvec may be the same buffer (of full local length). */
int jots_ctx_expand_true(jots_ctx_t c, const double* tvec, double* lvec, int loc);

/* ---- restart files: the VisItDataCollection of OutputManager::WriteRestartOutput (src/output_manager.cpp:15-21,89-95) and
   the load path of JOTSDriver (src/jots_driver.cpp:214-257).  <prefix>_<cycle:06d>.mfem_root (JSON), <prefix>_<cycle:06d>/
   mesh.000000 (MFEM mesh v1.0) and Temperature.000000 (H1 GridFunction, precision 16, values in MFEM's DOF numbering:
   vertices, edges, faces, interiors -- restated from MFEM 4.7, see restart.cpp).  Single-domain form: on several GPUs rank 0
   writes the whole field.  The driver shim honours Use_Restart / Restart_Prefix / Restart_Cycle / Restart_Freq.
   jots_mfem_dof_numbering: out[node of the space] = MFEM DOF number; jots_restart_read returns the mesh (caller destroys) and
   the values in MFEM numbering (call once with mfem_values == NULL for the count). */
int jots_mfem_dof_numbering(jots_mesh_t m, jots_space_t s, int64_t* out);
int jots_restart_write(const char* prefix, int cycle, double time, jots_mesh_t m, jots_space_t s, const double* u_nodes);
jots_mesh_t jots_restart_read(const char* prefix, int cycle, int* order, int* cycle_out, double* time, int64_t* n_values, double* mfem_values);

/* ---- ParaView output: the ParaViewDataCollection of OutputManager::WriteVizOutput (src/output_manager.cpp:23-28,81-87):
   <dir>/ParaView.pvd, <dir>/Cycle<cycle:06d>/data.pvtu + proc000000.vtu; one VTK_LAGRANGE_QUADRILATERAL / _HEXAHEDRON cell of the
   FE order per element with its own points (high-order output, levels of detail = order), inline base64 binary, Float64; the
   nodal fields are evaluated at the uniform lattice VTK's Lagrange cells use.  mode 0: new collection, 1: add this cycle to an
   existing .pvd, 2: restart mode (keep the entries before `time`).  The driver shim writes "Rank", "Temperature" and one field per
   material property of the input file at Visualization_Freq (dir "ParaView", or $JOTS_PARAVIEW_DIR).  Restated from the VTK file
   format / MFEM 4.7 (paraview.cpp).  jots_vtk_lagrange_order: lex_of_vtk[j] = lexicographic index of VTK node j. */
int jots_paraview_write(const char* dir, int cycle, double time, jots_mesh_t m, jots_space_t s, int n_fields, const char* const* names,
                        const double* const* nodal_values, int mode);
int jots_vtk_lagrange_order(int dim, int p, int32_t* lex_of_vtk);

/* Config (src/config_file.cpp:7-172) parsed on the host, as "key=value" lines (scalars, then "mat:<name>=a|b|..." and
   "bc:<attribute>=a|b|..."): lets the key set, defaults and parsing quirks be compared with the reference's without a
   device.  Returns the length needed (including the terminating 0), -1 on error. */
int64_t jots_config_dump(const char* ini_path, char* out, int64_t cap);

/* ---- driver shim: Config + the hot-path part of JOTSDriver::Run (src/jots_driver.cpp:568-725) --- */
jots_driver_t jots_driver_create(const char* ini_path, int device);
jots_driver_t jots_driver_create_on_mesh(const char* ini_path, jots_mesh_t mesh, int device);
/* collective over the communicator: every rank passes the same global mesh (or NULL: read Mesh_File) */
jots_driver_t jots_driver_create_dist(const char* ini_path, jots_mesh_t mesh, int device, jots_comm_t comm);
int jots_driver_copy_global_ids(jots_driver_t d, int64_t* global_id, int64_t* n_owned);
void jots_driver_destroy(jots_driver_t d);
int jots_driver_run(jots_driver_t d, int max_steps, int* steps_done);  /* max_steps < 0: to Max_Timesteps */
int64_t jots_driver_size(jots_driver_t d);
int jots_driver_get_u(jots_driver_t d, double* u, int loc);
int jots_driver_set_u(jots_driver_t d, const double* u, int loc);
double jots_driver_time(jots_driver_t d);
jots_ctx_t jots_driver_ctx(jots_driver_t d);      /* borrowed */
jots_op_t jots_driver_op(jots_driver_t d);        /* borrowed: the driver's JOTSIterator */
int jots_driver_step_count(jots_driver_t d);      /* it_num */
jots_space_t jots_driver_space(jots_driver_t d);  /* new handle, caller destroys */

#ifdef __cplusplus
}
#endif
#endif /* JOTS_B200_H */
