// Device context of the conduction solve: uploaded space tables, boundary-condition index sets,
// material models, form application and the BLAS-1 helpers the solvers use.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "diag_face_host.hpp"
#include "forms.hpp"
#include "mesh.hpp"
#include "vec.hpp"

namespace jots {

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
// apply_form was asked to carry the CG side stream with a form / vectors that cannot (callers probe with form_side_chunk first)
struct NoSideStream : std::runtime_error {
    using std::runtime_error::runtime_error;
};
void cuda_check(cudaError_t e, const char* what, const char* file, int line);
#define JOTS_CUDA(x) ::jots::cuda_check((x), #x, __FILE__, __LINE__)

// NVTX range for the tools (`ncu --nvtx --nvtx-include "jots:krylov/"` profiles only the kernels of the Krylov solves):
// step > newton > krylov, diagonal.  Two calls through a function pointer that is null when no tool is attached.
struct NvtxRange {
    explicit NvtxRange(const char* name);
    ~NvtxRange();
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// simple owning device array
template <class T>
struct DArr {
    T* p = nullptr;
    size_t n = 0;
    DArr() = default;
    DArr(const DArr&) = delete;
    DArr(DArr&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DArr& operator=(const DArr&) = delete;
    ~DArr() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    void alloc(size_t count) {
        if (count == n && p) return;
        release();
        if (count) JOTS_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
        n = count;
    }
    void upload(const std::vector<T>& h) {
        alloc(h.size());
        if (!h.empty()) JOTS_CUDA(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    }
};

enum { BC_ISOTHERMAL = 0, BC_HEATFLUX = 1, BC_SIN_ISOTHERMAL = 2, BC_SIN_HEATFLUX = 3,
       BC_PRECICE_ISOTHERMAL = 4, BC_PRECICE_HEATFLUX = 5 };   // PreciceBC: v[0] = default value, data per boundary node

// BC value objects (/root/reference/src/boundary_condition.hpp:14-105, .cpp:49-75)
struct BCSpec {
    int kind = BC_HEATFLUX;
    double v[4] = {0, 0, 0, 0};  // value | amp, angular freq, phase, vertical shift
    double value = 0.0;          // current coefficient value (UpdateCoeff)
    bool essential() const { return kind == BC_ISOTHERMAL || kind == BC_SIN_ISOTHERMAL || kind == BC_PRECICE_ISOTHERMAL; }
    bool constant() const { return kind == BC_ISOTHERMAL || kind == BC_HEATFLUX; }
    bool precice() const { return kind == BC_PRECICE_ISOTHERMAL || kind == BC_PRECICE_HEATFLUX; }
    void update(double t);
};

// one matrix-free operator = one element form + optional boundary-mass term + essential handling
struct FormSpec {
    int op = OP_LIN, cf = CF_CONST;
    double cm = 0, ck = 0, cb = 0, m0 = 1, a0 = 1, inv_rc = 1;
    bool mask_in = false, mask_out = false;
    bool diag_one = false;          // y[ess] = x[ess] (EliminateRowsCols / DIAG_ONE)
    const double* state = nullptr;  // frozen state z
    const double* coef_state = nullptr;  // properties evaluated on this vector instead of z (null: on z)
    double bdr_mass_scale = 0.0;    // + scale * int_bdr (g/(rho C))' phi_i phi_j  (0 = absent)
};

struct Stats {
    long long applies = 0, diag_assemblies = 0, krylov_iters = 0, newton_iters = 0, residual_evals = 0, steps = 0,
              kernel_launches = 0, dots = 0, halo_exchanges = 0, patch_applies = 0, fused_dots = 0, side_applies = 0;
    double last_linear_norm = 0, last_newton_norm = 0;
    int last_linear_converged = 1, last_newton_converged = 1, last_linear_iters = 0, last_newton_iters = 0;
};

struct Halo;  // multi-GPU interface exchange (halo.hpp)

class Ctx {
public:
    Ctx(const Space& s, int device);
    // one subdomain of a partitioned space; comm may be null (no exchange: testing only)
    Ctx(const struct Partition& p, int device, struct Comm* comm);
    ~Ctx();

    // --- static description
    int device = 0;
    cudaStream_t stream = nullptr;
    Space space;             // host copy (index sets, boundary tables)
    int64_t n = 0;           // local vector length (L-vector)
    int64_t n_owned = 0;     // leading entries owned by this rank (== n on one GPU)
    TablesRT tabs;
    MatSet mat;
    std::vector<BCSpec> bcs;             // index i <-> boundary attribute i+1
    std::vector<int> ess;                // sorted essential DOFs (GetEssentialTrueDofs)
    std::vector<std::vector<int>> bc_dofs;  // per BC: DOFs of its boundary elements
    Stats stats;
    const double* apply_skip = nullptr;   // device flag forwarded to the element kernels (FormArgs::skip)
    const char* last_kernel = "";     // element kernel of the last apply_form (launch.hpp: last_element_kernel)
    Halo* halo = nullptr;

    // --- device tables
    DArr<int> d_gather, d_gather_ess, d_bdr_gather, d_bdr_attr, d_ess, d_ess_owned;
    int n_ess_owned = 0;
    std::vector<DArr<int>> d_bc_dofs;
    // preCICE boundary data path (/root/reference/src/boundary_condition.cpp:77-206): per preCICE BC the nodal coefficient field
    // (coeff_gf: default value, boundary entries overwritten by the read data), the boundary elements of its attribute and
    // their nodes boundary element by boundary element (duplicates kept, the order of coords / read / write arrays)
    std::vector<DArr<double>> bc_field;
    std::vector<std::vector<int>> bc_faces, bc_node_dofs;
    std::vector<DArr<int>> d_bc_faces, d_bc_node_dofs;
    DArr<const double*> d_gnodal;       // per BC: bc_field pointer for preCICE heat-flux BCs, else null
    DArr<int> d_bdr_elem, d_bdr_face, d_bdr_node_loc;   // adjacent element, its local face, node places of every boundary element
    int64_t precice_bc_size(int bc) const { return (int64_t)bc_node_dofs.at(bc).size(); }
    void precice_bc_coords(int bc, double* xyz) const;                 // [size][dim]
    void precice_bc_set_read_data(int bc, const double* vals_host);    // PreciceBC::UpdateCoeff: SetSubVector(bdr_dof_indices, read)
    void precice_bc_write_data(int bc, const double* u, double* out_host);   // RetrieveWriteData: wall heat flux | temperatures
    void apply_dirichlet(int bc, double* u);                            // ProjectBdrCoefficient of BC `bc` (uniform value or field)
    DArr<double> d_geom, d_geomq, d_bdr_corner, d_gattr;
    DArr<unsigned char> d_essmark;
    int slab_variant = 7;             // JOTS_SLAB_VARIANT: 7 unique-pencil patch kernels where the mesh is a lattice (default), 3 gather-table kernel only
    DArr<int> d_patches;              // element patches of the unique-pencil kernel (patches.cpp); empty: not applicable
    int n_patches = 0;
    double patch_fill = 0.0;          // nelem / (patches * elements per patch)
    void build_patches();             // rebuilt with the essential set; used when the fill is at least JOTS_PATCH_MIN_FILL (0.6)
    int geom_stride = 10;
    bool geom_diag = false;   // diagonal metric on every element
    bool use_slab = true;     // JOTS_DISABLE_SLAB=1 forces the generic kernel
    DArr<double> d_scal, d_partial;
    DArr<double> mat_state;   // vector the nodal property polynomials were last evaluated on (UpdateMatProps)
    void set_mat_state(const double* u) { mat_state.alloc(n); copy(u, mat_state.p); }
    DArr<unsigned int> d_counter;
    DArr<unsigned int> d_sched;   // tile / exit counters of the patch kernel's dynamic scheduler (self-resetting)
    double* h_scal = nullptr;  // pinned

    // --- setup
    void set_materials(const Poly& rho, const Poly& C, const Poly& k);
    void set_bcs(const std::vector<BCSpec>& b);

    // --- forms
    // dot != nullptr: also accumulate (x, A x) over the owned rows into *dot when the launched kernel can (returns true; the
    // caller all-reduces); false = y only
    // side != nullptr (only after form_side_chunk(f) > 0, dot != nullptr and y ALREADY zero): the launch also carries the
    // conjugate-gradient loop's side stream, side->x += *side->alpha * side->d and side->zero = 0 (FormArgs::side_*)
    bool apply_form(const FormSpec& f, const double* x, double* y, double* dot = nullptr, const CgSide* side = nullptr);
    bool ess_rows_are_zero(const double* v);        // v[i] == 0 on every owned essential DOF (one small kernel + a synchronisation)
    int form_side_chunk(const FormSpec& f) const;   // entries per tile of the side stream for this form; 0 = not available
    void diag_form(const FormSpec& f, double* d);
    // b_i = sum_faces int (g / divisor) phi_i; state != null -> divisor rho_h(state) C_h(state)
    void neumann_vector(double div_const, const double* state, double* b, bool zero_first = true, double scale = 1.0,
                        bool mask_out = false);
    void upload_flux_values();   // g per attribute from the current BC values (essential -> 0)
    bool any_flux() const;

    // --- vectors
    DotScratch dot_scratch() { return DotScratch{d_partial.p, d_counter.p}; }
    double dot(const double* a, const double* b);
    void dots(int nv, const double* const* a, const double* const* b, double* out);
    void fetch_scalars(int nv, double* out);   // d_scal -> host (+ allreduce across ranks)
    void allreduce_scalars(double* d_vals, int nv);   // stream-ordered all-reduce of device scalars (no-op on one GPU)
    void copy_scalars(int nv, double* out);            // d_scal[0..nv) -> host, one synchronisation
    double norm(const double* a) ;
    double max_local(const double* a);   // rank-local maximum, as Vector::Max() in the reference
    double max_global(const double* a);  // maximum over all ranks (ncclMax): collective
    int rank() const;                    // 0 on one GPU
    void zero(double* y) { JOTS_CUDA(cudaMemsetAsync(y, 0, n * sizeof(double), stream)); }
    void copy(const double* x, double* y) { JOTS_CUDA(cudaMemcpyAsync(y, x, n * sizeof(double), cudaMemcpyDeviceToDevice, stream)); }
    void zero_ess(double* y) { v_set_idx(stream, (int)ess.size(), d_ess.p, 0.0, y); stats.kernel_launches += !ess.empty(); }
    void sync() { JOTS_CUDA(cudaStreamSynchronize(stream)); }
    // owner -> ghost broadcast of a local vector whose owned (leading n_owned) entries are valid: P of the reference's
    // ParFiniteElementSpace, done as zero-the-ghosts + interface sum-exchange.  No-op on one GPU.
    void expand_owned(double* v);
    // run on the caller's stream from now on: the context's own stream is drained and destroyed, the caller keeps
    // ownership of s (it is never destroyed here)
    void set_stream(cudaStream_t s);

    FormArgs make_args(const FormSpec& f, const double* x, double* y, bool for_diag) const;
private:
    void init(const Space& s, int dev);
    bool own_stream = false;
};

}  // namespace jots
