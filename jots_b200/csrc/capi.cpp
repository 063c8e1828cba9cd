// extern "C" boundary of libjots_b200.so (declarations and reference citations: include/jots_b200.h)
#include <cstdlib>
#include "../../include/jots_b200.h"

#include <algorithm>
#include <cstring>
#include <sstream>
#include <exception>
#include <string>

#include "driver.hpp"
#include "halo.hpp"
#include "launch.hpp"
#include "paraview.hpp"
#include "restart.hpp"

using namespace jots;

struct jots_mesh_s { Mesh m; };
struct jots_space_s { Space s; };
struct jots_ctx_s { std::unique_ptr<Ctx> own; Ctx* c = nullptr; };
struct jots_op_s {
    Ctx* c = nullptr;
    int sim_type = 0;
    std::unique_ptr<JOTSIterator> own;
    JOTSIterator* it = nullptr;
    DArr<double> un;  // u_n copy for residual building blocks
};
struct jots_comm_s { Comm* c = nullptr; };
struct jots_part_s { Partition p; };
struct jots_driver_s { std::unique_ptr<JOTSDriver> d; jots_ctx_s ctx_view; jots_op_s op_view; };

static thread_local std::string g_err;
static int fail(const std::exception& e) { g_err = e.what(); return 1; }
#define JOTS_TRY try {
#define JOTS_CATCH_INT } catch (const std::exception& e) { return fail(e); } catch (...) { g_err = "unknown error"; return 1; }
#define JOTS_CATCH_PTR } catch (const std::exception& e) { fail(e); return nullptr; } catch (...) { g_err = "unknown error"; return nullptr; }

namespace {
// stage a vector argument on the device
struct In {
    const double* p = nullptr;
    DArr<double> tmp;
    In(Ctx* c, const double* v, int loc) {
        if (!v) return;
        if (loc == JOTS_DEVICE) { p = v; return; }
        tmp.alloc(c->n);
        JOTS_CUDA(cudaMemcpyAsync(tmp.p, v, c->n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        p = tmp.p;
    }
};
struct Out {
    double* p = nullptr;
    double* host = nullptr;
    Ctx* c;
    DArr<double> tmp;
    Out(Ctx* ctx, double* v, int loc, bool load = false) : c(ctx) {
        if (loc == JOTS_DEVICE) { p = v; return; }
        host = v;
        tmp.alloc(c->n);
        p = tmp.p;
        if (load) JOTS_CUDA(cudaMemcpyAsync(p, v, c->n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    void finish() {
        if (host) JOTS_CUDA(cudaMemcpyAsync(host, p, c->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        JOTS_CUDA(cudaStreamSynchronize(c->stream));
    }
};
}  // namespace

extern "C" {

const char* jots_last_error(void) { return g_err.c_str(); }
const char* jots_version(void) { return "jots_b200 0.1 (sm_100a)"; }
int jots_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ---- mesh -------------------------------------------------------------------------------------
jots_mesh_t jots_mesh_read(const char* path) {
    JOTS_TRY auto* h = new jots_mesh_s; h->m = read_mfem_mesh(path); return h; JOTS_CATCH_PTR
}
jots_mesh_t jots_mesh_brick(int64_t nx, int64_t ny, int64_t nz, double lx, double ly, double lz) {
    JOTS_TRY
    if (nx < 1 || ny < 1 || nz < 1) throw std::runtime_error("brick needs at least one element per direction");
    auto* h = new jots_mesh_s; h->m = make_brick(nx, ny, nz, lx, ly, lz); return h;
    JOTS_CATCH_PTR
}
jots_mesh_t jots_mesh_rect(int64_t nx, int64_t ny, double lx, double ly, double x0, double y0) {
    JOTS_TRY auto* h = new jots_mesh_s; h->m = make_rect(nx, ny, lx, ly, x0, y0); return h; JOTS_CATCH_PTR
}
jots_mesh_t jots_mesh_refine(jots_mesh_t m) {
    JOTS_TRY auto* h = new jots_mesh_s; h->m = uniform_refine(m->m); return h; JOTS_CATCH_PTR
}
void jots_mesh_destroy(jots_mesh_t m) { delete m; }
int jots_mesh_info(jots_mesh_t m, int* dim, int64_t* ne, int64_t* nb, int64_t* nv) {
    JOTS_TRY
    if (dim) *dim = m->m.dim;
    if (ne) *ne = m->m.n_elem();
    if (nb) *nb = m->m.n_bdr();
    if (nv) *nv = m->m.n_verts();
    return 0;
    JOTS_CATCH_INT
}
int jots_mesh_copy(jots_mesh_t m, double* verts, int64_t* elems, int64_t* bdr, int32_t* bdr_attr) {
    JOTS_TRY
    const Mesh& M = m->m;
    if (verts) std::memcpy(verts, M.verts.data(), M.verts.size() * sizeof(double));
    if (elems) std::memcpy(elems, M.elems.data(), M.elems.size() * sizeof(int64_t));
    if (bdr) std::memcpy(bdr, M.bdr.data(), M.bdr.size() * sizeof(int64_t));
    if (bdr_attr) for (size_t i = 0; i < M.bdr_attr.size(); ++i) bdr_attr[i] = M.bdr_attr[i];
    return 0;
    JOTS_CATCH_INT
}
int jots_mesh_set_vertices(jots_mesh_t m, const double* verts) {
    JOTS_TRY
    if (!verts) throw std::runtime_error("jots_mesh_set_vertices: null coordinates");
    std::memcpy(m->m.verts.data(), verts, m->m.verts.size() * sizeof(double));
    return 0;
    JOTS_CATCH_INT
}

// ---- space ------------------------------------------------------------------------------------
jots_space_t jots_space_create(jots_mesh_t m, int order) {
    JOTS_TRY auto* h = new jots_space_s; h->s = make_space(m->m, order); return h; JOTS_CATCH_PTR
}
// host only: the adjacency the preCICE write path needs (Ctx::set_bcs), exposed so that the index contract is testable without a GPU
int jots_space_boundary_adjacency(jots_space_t sp, int32_t* bdr_elem, int32_t* bdr_face, int32_t* node_loc) {
    JOTS_TRY
    Space s = sp->s;      // a copy: the handle keeps what the caller gave it
    fill_boundary_adjacency(s);
    std::vector<int> nl;
    boundary_node_locals(s, nl);
    for (int64_t b = 0; b < s.nbdr; ++b) {
        if (bdr_elem) bdr_elem[b] = (int32_t)s.bdr_elem[b];
        if (bdr_face) bdr_face[b] = (int32_t)s.bdr_face[b];
    }
    if (node_loc) std::copy(nl.begin(), nl.end(), node_loc);
    return 0;
    JOTS_CATCH_INT
}
jots_space_t jots_space_create_from_tables(int dim, int order, int64_t n_elem, int64_t n_dof, const int32_t* gather,
                                           const double* elem_corners, int64_t n_bdr, const int32_t* bdr_gather,
                                           const double* bdr_corners, const int32_t* bdr_attr) {
    JOTS_TRY
    if (dim != 2 && dim != 3) throw std::runtime_error("only 2-D / 3-D spaces are supported");
    if (order < 1 || order > 4) throw std::runtime_error("FE_Order must be 1..4");
    auto* h = new jots_space_s;
    Space& S = h->s;
    S.dim = dim; S.p = order; S.D = order + 1; S.Q = dim == 3 ? order + 2 : order + 1;
    S.nloc = 1; for (int a = 0; a < dim; ++a) S.nloc *= S.D;
    S.nfl = S.nloc / S.D;
    S.nelem = n_elem; S.ndof = n_dof; S.nbdr = n_bdr;
    const int nv = 1 << dim, nfv = 1 << (dim - 1);
    S.gather.assign(gather, gather + n_elem * S.nloc);
    for (int g : S.gather) if (g < 0 || g >= n_dof) { delete h; throw std::runtime_error("gather table entry out of range"); }
    S.elem_corner.assign(elem_corners, elem_corners + n_elem * nv * dim);
    S.bdr_gather.assign(bdr_gather, bdr_gather + n_bdr * S.nfl);
    S.bdr_corner.assign(bdr_corners, bdr_corners + n_bdr * nfv * dim);
    S.bdr_attr.assign(bdr_attr, bdr_attr + n_bdr);
    S.bdr_elem.assign(n_bdr, -1); S.bdr_face.assign(n_bdr, -1);
    S.attributes = S.bdr_attr;
    std::sort(S.attributes.begin(), S.attributes.end());
    S.attributes.erase(std::unique(S.attributes.begin(), S.attributes.end()), S.attributes.end());
    {
        // GL nodes for node_coordinates()
        Mesh dummy = dim == 3 ? make_brick(1, 1, 1, 1, 1, 1) : make_rect(1, 1, 1, 1, 0, 0);
        S.nodes = make_space(dummy, order).nodes;
    }
    return h;
    JOTS_CATCH_PTR
}
int jots_space_copy_corners(jots_space_t s, double* elem_corners, double* bdr_corners) {
    JOTS_TRY
    if (elem_corners) std::memcpy(elem_corners, s->s.elem_corner.data(), s->s.elem_corner.size() * sizeof(double));
    if (bdr_corners) std::memcpy(bdr_corners, s->s.bdr_corner.data(), s->s.bdr_corner.size() * sizeof(double));
    return 0;
    JOTS_CATCH_INT
}
void jots_space_destroy(jots_space_t s) { delete s; }
int jots_space_info(jots_space_t s, int* dim, int* order, int64_t* ne, int64_t* nd, int64_t* nb, int* dpe, int* dpf) {
    JOTS_TRY
    const Space& S = s->s;
    if (dim) *dim = S.dim;
    if (order) *order = S.p;
    if (ne) *ne = S.nelem;
    if (nd) *nd = S.ndof;
    if (nb) *nb = S.nbdr;
    if (dpe) *dpe = S.nloc;
    if (dpf) *dpf = S.nfl;
    return 0;
    JOTS_CATCH_INT
}
int jots_space_copy_gather(jots_space_t s, int32_t* gather) {
    JOTS_TRY std::memcpy(gather, s->s.gather.data(), s->s.gather.size() * sizeof(int)); return 0; JOTS_CATCH_INT
}
int jots_space_copy_bdr(jots_space_t s, int32_t* bg, int32_t* ba, int64_t* be, int32_t* bf) {
    JOTS_TRY
    const Space& S = s->s;
    if (bg) std::memcpy(bg, S.bdr_gather.data(), S.bdr_gather.size() * sizeof(int));
    if (ba) std::memcpy(ba, S.bdr_attr.data(), S.bdr_attr.size() * sizeof(int));
    if (be) std::memcpy(be, S.bdr_elem.data(), S.bdr_elem.size() * sizeof(int64_t));
    if (bf) std::memcpy(bf, S.bdr_face.data(), S.bdr_face.size() * sizeof(int));
    return 0;
    JOTS_CATCH_INT
}
int jots_space_copy_nodes(jots_space_t s, double* xyz) {
    JOTS_TRY
    std::vector<double> out;
    s->s.node_coordinates(out);
    std::memcpy(xyz, out.data(), out.size() * sizeof(double));
    return 0;
    JOTS_CATCH_INT
}
int64_t jots_space_essential_dofs(jots_space_t s, const int32_t* attrs, int n_attrs, int32_t* out) {
    try {
        std::vector<int> a(attrs, attrs + n_attrs), e;
        s->s.essential_dofs(a, e);
        if (out) std::memcpy(out, e.data(), e.size() * sizeof(int));
        return (int64_t)e.size();
    } catch (const std::exception& ex) { fail(ex); return -1; }
}

int64_t jots_space_patches(jots_space_t s, const int32_t* ess, int64_t n_ess, int px, int py, int* record_ints, double* fill, int32_t* out) {
    try {
        std::vector<int> e(ess, ess + n_ess);
        std::vector<int> tab;
        const int64_t np = make_patches(s->s, e, px, py, tab, fill);
        if (record_ints) *record_ints = np > 0 ? (int)(tab.size() / (size_t)np) : 0;
        if (out && np > 0) std::memcpy(out, tab.data(), tab.size() * sizeof(int));
        return np;
    } catch (const std::exception& ex) { fail(ex); return -1; }
}

int jots_patch_shape(int order, int* px, int* py, uint32_t* items) {
    int x = 0, y = 0;
    patch_shape(order, x, y);
    if (px) *px = x;
    if (py) *py = y;
    return (items && x > 0) ? patch_items(order, items) : 0;
}

int jots_h1_dof_map(int dim, int order, int32_t* map) {
    JOTS_TRY
    std::vector<int> m;
    h1_dof_map(dim, order, m);
    for (size_t i = 0; i < m.size(); ++i) map[i] = m[i];
    return 0;
    JOTS_CATCH_INT
}
int jots_gather_from_native(int dim, int order, int64_t n_elem, const int32_t* native, int32_t* lexicographic) {
    JOTS_TRY
    std::vector<int> m;
    h1_dof_map(dim, order, m);
    const size_t nl = m.size();
    for (int64_t e = 0; e < n_elem; ++e)
        for (size_t l = 0; l < nl; ++l) {
            const int32_t d = native[(size_t)e * nl + m[l]];
            lexicographic[(size_t)e * nl + l] = d < 0 ? -1 - d : d;      // H1 DOFs carry no sign: -1-d encodes d
        }
    return 0;
    JOTS_CATCH_INT
}

// ---- restart files (host only) ----------------------------------------------------------------------
int jots_mfem_dof_numbering(jots_mesh_t m, jots_space_t s, int64_t* out) {
    JOTS_TRY
    std::vector<int64_t> num;
    mfem_dof_numbering(m->m, s->s.p, s->s.gather, s->s.ndof, num);
    std::memcpy(out, num.data(), num.size() * sizeof(int64_t));
    return 0;
    JOTS_CATCH_INT
}
int jots_restart_write(const char* prefix, int cycle, double time, jots_mesh_t m, jots_space_t s, const double* u_nodes) {
    JOTS_TRY write_restart(prefix, cycle, time, m->m, s->s, u_nodes); return 0; JOTS_CATCH_INT
}
int jots_paraview_write(const char* dir, int cycle, double time, jots_mesh_t m, jots_space_t s, int n_fields, const char* const* names,
                        const double* const* nodal_values, int mode) {
    JOTS_TRY
    ParaViewCollection pc;
    if (dir && dir[0]) pc.dir = dir;
    pc.append = (mode == 1);
    pc.restart_mode = (mode == 2);
    std::vector<VizField> fields;
    for (int i = 0; i < n_fields; ++i) fields.push_back({names[i], nodal_values[i]});
    pc.save(cycle, time, s->s, m ? m->m.elem_attr : std::vector<int>(), fields);
    return 0;
    JOTS_CATCH_INT
}
int jots_vtk_lagrange_order(int dim, int p, int32_t* lex_of_vtk) {
    JOTS_TRY
    std::vector<int> v;
    vtk_lagrange_order(dim, p, v);
    for (size_t i = 0; i < v.size(); ++i) lex_of_vtk[i] = v[i];
    return 0;
    JOTS_CATCH_INT
}
jots_mesh_t jots_restart_read(const char* prefix, int cycle, int* order, int* cycle_out, double* time, int64_t* n_values, double* mfem_values) {
    JOTS_TRY
    RestartData d = read_restart(prefix, cycle);
    if (order) *order = d.order;
    if (cycle_out) *cycle_out = d.cycle;
    if (time) *time = d.time;
    if (n_values) *n_values = (int64_t)d.mfem_values.size();
    if (mfem_values) std::memcpy(mfem_values, d.mfem_values.data(), d.mfem_values.size() * sizeof(double));
    auto* h = new jots_mesh_s;
    h->m = std::move(d.mesh);
    return h;
    JOTS_CATCH_PTR
}

// ---- config (host only) -----------------------------------------------------------------------
int64_t jots_config_dump(const char* ini_path, char* out, int64_t cap) {
    try {
        Config cfg(ini_path);
        std::ostringstream os;
        os.precision(17);
        os << "sim_type=" << cfg.sim_type << "\nmesh_file=" << cfg.mesh_file << "\nuse_restart=" << (int)cfg.use_restart
           << "\nfe_order=" << cfg.fe_order << "\nserial_refine=" << cfg.serial_refine << "\nparallel_refine=" << cfg.parallel_refine
           << "\ninitial_temp=" << cfg.initial_temp << "\nwith_precice=" << (int)cfg.with_precice
           << "\nusing_time_integration=" << (int)cfg.using_time_integration << "\ntime_scheme=" << cfg.time_scheme << "\ndt=" << cfg.dt
           << "\nmax_timesteps=" << cfg.max_timesteps << "\nsolver=" << cfg.solver << "\nprec=" << cfg.prec << "\nabs_tol=" << cfg.abs_tol
           << "\nrel_tol=" << cfg.rel_tol << "\nmax_iter=" << cfg.max_iter << "\nusing_newton=" << (int)cfg.using_newton
           << "\nnewton_max_iter=" << cfg.newton_max_iter << "\nnewton_abs_tol=" << cfg.newton_abs_tol << "\nnewton_rel_tol=" << cfg.newton_rel_tol << "\n";
        for (auto& kv : cfg.mat_props) {
            os << "mat:" << kv.first << "=";
            for (size_t i = 0; i < kv.second.size(); ++i) os << (i ? "|" : "") << kv.second[i];
            os << "\n";
        }
        for (auto& kv : cfg.bcs) {
            os << "bc:" << kv.first << "=";
            for (size_t i = 0; i < kv.second.size(); ++i) os << (i ? "|" : "") << kv.second[i];
            os << "\n";
        }
        const std::string s = os.str();
        if (out && cap > 0) {
            const size_t nc = std::min<size_t>(s.size(), (size_t)cap - 1);
            std::memcpy(out, s.data(), nc);
            out[nc] = 0;
        }
        return (int64_t)s.size() + 1;
    } catch (const std::exception& ex) { fail(ex); return -1; }
}

// ---- context ----------------------------------------------------------------------------------
jots_ctx_t jots_ctx_create(jots_space_t s, int device) {
    JOTS_TRY
    auto* h = new jots_ctx_s;
    h->own.reset(new Ctx(s->s, device));
    h->c = h->own.get();
    return h;
    JOTS_CATCH_PTR
}
void jots_ctx_destroy(jots_ctx_t c) { delete c; }
int64_t jots_ctx_size(jots_ctx_t c) { return c->c->n; }
int jots_ctx_set_stream(jots_ctx_t c, void* s) {
    JOTS_TRY
    c->c->set_stream((cudaStream_t)s);
    return 0;
    JOTS_CATCH_INT
}
int jots_ctx_expand_true(jots_ctx_t c, const double* tvec, double* lvec, int loc) {
    JOTS_TRY
    Ctx* x = c->c;
    Out o(x, lvec, loc);
    const cudaMemcpyKind kind = loc == JOTS_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (!(loc == JOTS_DEVICE && tvec == lvec))
        JOTS_CUDA(cudaMemcpyAsync(o.p, tvec, (size_t)x->n_owned * sizeof(double), kind, x->stream));
    x->expand_owned(o.p);
    o.finish();
    return 0;
    JOTS_CATCH_INT
}
int jots_ctx_sync(jots_ctx_t c) { JOTS_TRY c->c->sync(); return 0; JOTS_CATCH_INT }
int jots_ctx_set_material(jots_ctx_t c, int which, int model, const double* coeffs, int n) {
    JOTS_TRY
    Poly p; p.n = 0; for (double& v : p.c) v = 0.0;
    if (model == JOTS_UNIFORM) { if (n < 1) throw std::runtime_error("Uniform property needs a value"); p.n = 1; p.c[0] = coeffs[0]; }
    else if (model == JOTS_POLYNOMIAL) {
        if (n < 2) throw std::runtime_error("Polynomial property needs >= 2 coefficients");
        if (n > MAX_POLY) throw std::runtime_error("Polynomial property with more than 8 coefficients is not supported");
        p.n = n; for (int i = 0; i < n; ++i) p.c[i] = coeffs[i];
    } else throw std::runtime_error("Unknown/Invalid material model specified");
    MatSet m = c->c->mat;
    if (which == JOTS_DENSITY) m.rho = p; else if (which == JOTS_SPECIFIC_HEAT) m.C = p;
    else if (which == JOTS_THERMAL_CONDUCTIVITY) m.k = p; else throw std::runtime_error("unknown material property id");
    c->c->set_materials(m.rho, m.C, m.k);
    return 0;
    JOTS_CATCH_INT
}
int jots_ctx_set_bcs(jots_ctx_t c, int n_bc, const int32_t* kinds, const double* params) {
    JOTS_TRY
    if ((size_t)n_bc != c->c->space.attributes.size()) throw std::runtime_error("Input file BC count and mesh BC counts do not match.");
    std::vector<BCSpec> b(n_bc);
    for (int i = 0; i < n_bc; ++i) {
        if (kinds[i] < 0 || kinds[i] > BC_PRECICE_HEATFLUX) throw std::runtime_error("Invalid/Unknown boundary condition specified");
        b[i].kind = kinds[i];
        for (int j = 0; j < 4; ++j) b[i].v[j] = params[4 * i + j];
        b[i].update(0.0);
    }
    c->c->set_bcs(b);
    return 0;
    JOTS_CATCH_INT
}
int64_t jots_ctx_essential_dofs(jots_ctx_t c, int32_t* out) {
    if (out) std::memcpy(out, c->c->ess.data(), c->c->ess.size() * sizeof(int));
    return (int64_t)c->c->ess.size();
}
int jots_ctx_update_bcs(jots_ctx_t c, double time) {
    JOTS_TRY for (auto& b : c->c->bcs) b.update(time); return 0; JOTS_CATCH_INT
}
int jots_ctx_apply_dirichlet(jots_ctx_t c, double* u, int loc, int only_td) {
    JOTS_TRY
    Ctx* x = c->c;
    Out o(x, u, loc, true);
    for (size_t i = 0; i < x->bcs.size(); ++i)
        if (x->bcs[i].essential() && (!only_td || !x->bcs[i].constant())) x->apply_dirichlet((int)i, o.p);
    o.finish();
    return 0;
    JOTS_CATCH_INT
}
int64_t jots_ctx_precice_bc_size(jots_ctx_t c, int bc) {
    Ctx* x = c->c;
    if (bc < 0 || bc >= (int)x->bcs.size() || !x->bcs[bc].precice()) return -1;
    return x->precice_bc_size(bc);
}
int jots_ctx_precice_bc_nodes(jots_ctx_t c, int bc, double* coords, int32_t* dofs) {
    JOTS_TRY
    Ctx* x = c->c;
    if (coords) x->precice_bc_coords(bc, coords);
    if (dofs) { const std::vector<int>& nd = x->bc_node_dofs.at(bc); for (size_t i = 0; i < nd.size(); ++i) dofs[i] = nd[i]; }
    return 0;
    JOTS_CATCH_INT
}
int jots_ctx_precice_bc_set_read_data(jots_ctx_t c, int bc, const double* values) {
    JOTS_TRY c->c->precice_bc_set_read_data(bc, values); return 0; JOTS_CATCH_INT
}
int jots_ctx_precice_bc_write_data(jots_ctx_t c, int bc, const double* u, int loc, double* out) {
    JOTS_TRY In i(c->c, u, loc); c->c->precice_bc_write_data(bc, i.p, out); return 0; JOTS_CATCH_INT
}
int jots_ctx_set_material_state(jots_ctx_t c, const double* u, int loc) {
    JOTS_TRY In i(c->c, u, loc); c->c->set_mat_state(i.p); c->c->sync(); return 0; JOTS_CATCH_INT
}
int jots_ctx_material_field(jots_ctx_t c, int which, int order, double* out, int loc) {
    JOTS_TRY
    Ctx* x = c->c;
    if (which < 0 || which > 2 || order < 0 || order > 2) throw std::runtime_error("jots_ctx_material_field: which in {0,1,2} (rho, C, k), order in {0,1,2}");
    const Poly& P = which == 0 ? x->mat.rho : which == 1 ? x->mat.C : x->mat.k;
    if (P.n > 1 && x->mat_state.n != (size_t)x->n)
        throw std::runtime_error("jots_ctx_material_field: temperature-dependent property but no material state set (jots_ctx_set_material_state)");
    Out o(x, out, loc);
    v_poly_field(x->stream, x->n, P, order, P.n > 1 ? x->mat_state.p : nullptr, o.p);
    ++x->stats.kernel_launches;
    o.finish();
    return 0;
    JOTS_CATCH_INT
}
int jots_ctx_stats(jots_ctx_t c, int64_t out[12], double norms[2]) {
    const Stats& s = c->c->stats;
    out[0] = s.applies; out[1] = s.diag_assemblies; out[2] = s.krylov_iters; out[3] = s.newton_iters;
    out[4] = s.residual_evals; out[5] = s.steps; out[6] = s.kernel_launches; out[7] = s.dots;
    out[8] = s.last_linear_iters; out[9] = s.last_linear_converged; out[10] = s.last_newton_iters; out[11] = s.last_newton_converged;
    if (norms) { norms[0] = s.last_linear_norm; norms[1] = s.last_newton_norm; }
    return 0;
}
int jots_ctx_patch_info(jots_ctx_t c, int64_t* n_patches, double* fill, int64_t* applies) {
    if (n_patches) *n_patches = c->c->n_patches;
    if (fill) *fill = c->c->patch_fill;
    if (applies) *applies = c->c->stats.patch_applies;
    return 0;
}
const char* jots_ctx_last_kernel(jots_ctx_t c) { return c->c->last_kernel; }
int jots_ctx_fused_dots(jots_ctx_t c, int64_t* n) { if (n) *n = c->c->stats.fused_dots; return 0; }
int jots_ctx_side_applies(jots_ctx_t c, int64_t* n) { if (n) *n = c->c->stats.side_applies; return 0; }
int jots_ctx_reset_stats(jots_ctx_t c) { c->c->stats = Stats(); return 0; }

// ---- operators --------------------------------------------------------------------------------
jots_op_t jots_op_create(jots_ctx_t c, int sim_type, double dt, int scheme, int solver, int prec, double rel, double abs_, int max_iter,
                         double nrel, double nabs, int nmax) {
    JOTS_TRY
    SolverSpec ls; ls.solver = solver; ls.prec = prec; ls.rel_tol = rel; ls.abs_tol = abs_; ls.max_iter = max_iter;
    NewtonSpec nw; nw.rel_tol = nrel; nw.abs_tol = nabs; nw.max_iter = nmax;
    if (solver < 0 || solver > 2) throw std::runtime_error("unknown solver id");
    if (prec < 0 || prec > 1) throw std::runtime_error("unknown preconditioner id");
    auto* h = new jots_op_s;
    h->c = c->c; h->sim_type = sim_type;
    if (!c->c->mat_state.p) { c->c->mat_state.alloc(c->c->n); c->c->zero(c->c->mat_state.p); }
    if (sim_type == JOTS_LINEARIZED_UNSTEADY) h->own.reset(new LinearConductionOperator(c->c, ls, dt, scheme));
    else if (sim_type == JOTS_NONLINEAR_UNSTEADY) h->own.reset(new NonlinearConductionOperator(c->c, ls, nw, dt, scheme));
    else if (sim_type == JOTS_STEADY) h->own.reset(new SteadyConductionOperator(c->c, ls, nw));
    else { delete h; throw std::runtime_error("unknown simulation type id"); }
    h->it = h->own.get();
    return h;
    JOTS_CATCH_PTR
}
void jots_op_destroy(jots_op_t op) { delete op; }
int jots_op_set_print_level(jots_op_t op, int linear_mask, int newton_mask) {
    JOTS_TRY op->it->set_print_level(linear_mask, newton_mask); return 0; JOTS_CATCH_INT
}
int jots_print_level_mask(const char* labels) {
    try {
        std::vector<std::string> v;
        std::stringstream ss(labels ? labels : "");
        std::string t;
        while (std::getline(ss, t, ',')) {
            const size_t a = t.find_first_not_of(" \t"), b = t.find_last_not_of(" \t");
            if (a != std::string::npos) v.push_back(t.substr(a, b - a + 1));
        }
        return create_print_level(v);
    } catch (const std::exception& ex) { fail(ex); return -1; }
}
int jots_op_iterate(jots_op_t op, double* u, int loc, double* time) {
    JOTS_TRY
    Out o(op->c, u, loc, true);
    if (time) op->it->time = *time;
    op->it->Iterate(o.p);
    if (time) *time = op->it->time;
    o.finish();
    return 0;
    JOTS_CATCH_INT
}

namespace {
struct OpAccess : JOTSIterator {  // reach the protected Mult / ImplicitSolve through the base
    static void implicit(JOTSIterator* it, double dt, const double* u, double* k) { (it->*(&OpAccess::ImplicitSolve))(dt, u, k); }
    static void mult(JOTSIterator* it, const double* u, double* k) { (it->*(&OpAccess::Mult))(u, k); }
};
}  // namespace

int jots_op_implicit_solve(jots_op_t op, double dt, const double* u, double* k, int loc) {
    JOTS_TRY
    In i(op->c, u, loc); Out o(op->c, k, loc);
    OpAccess::implicit(op->it, dt, i.p, o.p);
    o.finish();
    return 0;
    JOTS_CATCH_INT
}
int jots_op_mult(jots_op_t op, const double* u, double* k, int loc) {
    JOTS_TRY
    In i(op->c, u, loc); Out o(op->c, k, loc);
    OpAccess::mult(op->it, i.p, o.p);
    o.finish();
    return 0;
    JOTS_CATCH_INT
}
int jots_op_process_mat_prop_update(jots_op_t op, int which) { JOTS_TRY op->it->ProcessMatPropUpdate(which); return 0; JOTS_CATCH_INT }
int jots_op_update_neumann(jots_op_t op) { JOTS_TRY op->it->UpdateNeumann(); return 0; JOTS_CATCH_INT }
int jots_op_get_neumann(jots_op_t op, double* b, int loc) {
    JOTS_TRY Out o(op->c, b, loc); op->c->copy(op->it->b_vec.p, o.p); o.finish(); return 0; JOTS_CATCH_INT
}

// resolve a building-block form of an operator into something to run
// side: the launch also carries the conjugate-gradient side stream (y already zero; throws when the form's kernel has none)
static void run_form(jots_op_t op, int form, const double* x, const double* state, double* y, bool diag, double* d_dot = nullptr, bool* fused = nullptr,
                     const CgSide* side = nullptr) {
    Ctx* c = op->c;
    auto apply = [&](const FormSpec& f) { const bool fu = c->apply_form(f, x, y, d_dot, side); if (fused) *fused = fu; };
    auto no_dot = [&] { if (d_dot) throw std::runtime_error("this form is not a single bilinear form: no (x, A x)"); };
    if (op->sim_type == JOTS_LINEARIZED_UNSTEADY) {
        auto* L = static_cast<LinearConductionOperator*>(op->it);
        if (state) c->set_mat_state(state);
        FormSpec f;
        if (form == JOTS_FORM_MASS) f = L->M_spec();
        else if (form == JOTS_FORM_STIFFNESS) f = L->K_spec();
        else if (form == JOTS_FORM_SYSTEM) f = L->A_spec();
        else if (form == JOTS_FORM_RESIDUAL && !diag) { no_dot(); L->CalculateRHS(x, y); return; }
        else throw std::runtime_error("form not available for this operator");
        if (diag) c->diag_form(f, y); else apply(f);
    } else if (op->sim_type == JOTS_NONLINEAR_UNSTEADY) {
        auto* N = static_cast<NonlinearConductionOperator*>(op->it);
        if (form == JOTS_FORM_SYSTEM) {
            if (!state) throw std::runtime_error("the nonlinear Jacobian needs a state");
            FormSpec f = N->jacobian_spec(state);
            if (diag) c->diag_form(f, y); else apply(f);
        } else if (form == JOTS_FORM_RESIDUAL && !diag) {
            if (!state) throw std::runtime_error("the nonlinear residual needs u_n as state");
            no_dot();
            // R(k = x) with u_n = state
            op->un.alloc(c->n);
            c->copy(state, op->un.p);
            N->set_un(op->un.p);
            N->new_state(x);
            N->mult(x, y);
        } else if (form == JOTS_FORM_MASS) {
            FormSpec f;
            f.op = OP_LIN; f.cf = CF_CONST; f.cm = 1.0; f.ck = 0.0; f.m0 = 1.0; f.a0 = 0.0;
            f.mask_in = f.mask_out = f.diag_one = true;
            if (diag) c->diag_form(f, y); else apply(f);
        } else throw std::runtime_error("form not available for this operator");
    } else {
        auto* S = static_cast<SteadyConductionOperator*>(op->it);
        if (form == JOTS_FORM_SYSTEM) {
            if (!state) throw std::runtime_error("the steady Jacobian needs a state");
            FormOp* J = static_cast<FormOp*>(S->gradient(state));
            if (diag) J->assemble_diag(y); else apply(J->f);
        } else if (form == JOTS_FORM_RESIDUAL && !diag) { no_dot(); S->mult(x, y); }
        else throw std::runtime_error("form not available for this operator");
    }
}

int jots_op_form_apply(jots_op_t op, int form, const double* x, const double* state, double* y, int loc) {
    JOTS_TRY
    In ix(op->c, x, loc), is(op->c, state, loc); Out o(op->c, y, loc);
    run_form(op, form, ix.p, is.p, o.p, false);
    o.finish();
    return 0;
    JOTS_CATCH_INT
}
int jots_op_form_apply_dot(jots_op_t op, int form, const double* x, const double* state, double* y, int loc, double* dot, int* fused) {
    JOTS_TRY
    Ctx* c = op->c;
    In ix(c, x, loc), is(c, state, loc); Out o(c, y, loc);
    DArr<double> d_dot;
    d_dot.alloc(1);
    JOTS_CUDA(cudaMemsetAsync(d_dot.p, 0, sizeof(double), c->stream));
    bool fu = false;
    run_form(op, form, ix.p, is.p, o.p, false, d_dot.p, &fu);
    double v = 0.0;
    if (fu) {
        c->allreduce_scalars(d_dot.p, 1);
        JOTS_CUDA(cudaMemcpyAsync(&v, d_dot.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        c->sync();
    } else v = c->dot(ix.p, o.p);
    if (dot) *dot = v;
    if (fused) *fused = fu ? 1 : 0;
    o.finish();
    return 0;
    JOTS_CATCH_INT
}
int jots_op_form_apply_side(jots_op_t op, int form, const double* x, const double* state, double* y, int loc, double* dot,
                            double alpha, const double* side_d, double* side_x, double* side_zero, int* carried) {
    JOTS_TRY
    Ctx* c = op->c;
    if (loc != JOTS_DEVICE) throw std::runtime_error("jots_op_form_apply_side: device vectors only");
    if (carried) *carried = 0;
    DArr<double> d_s;
    d_s.alloc(2);
    const double hs[2] = {0.0, alpha};
    JOTS_CUDA(cudaMemcpyAsync(d_s.p, hs, sizeof(hs), cudaMemcpyHostToDevice, c->stream));
    c->zero(y);
    const CgSide side{side_x, side_d, d_s.p + 1, side_zero, 0};
    bool fu = false;
    try { run_form(op, form, x, state, y, false, d_s.p, &fu, &side); }
    catch (const NoSideStream&) { c->sync(); return 0; }      // *carried stays 0: this form has no side stream
    c->allreduce_scalars(d_s.p, 1);
    double v = 0.0;
    JOTS_CUDA(cudaMemcpyAsync(&v, d_s.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    c->sync();
    if (dot) *dot = v;
    if (carried) *carried = 1;
    return 0;
    JOTS_CATCH_INT
}
int jots_op_form_diagonal(jots_op_t op, int form, const double* state, double* d, int loc) {
    JOTS_TRY
    In is(op->c, state, loc); Out o(op->c, d, loc);
    run_form(op, form, nullptr, is.p, o.p, true);
    o.finish();
    return 0;
    JOTS_CATCH_INT
}
int jots_op_form_time(jots_op_t op, int form, const double* x, const double* state, double* y, int reps, double* ms) {
    JOTS_TRY
    Ctx* c = op->c;
    cudaEvent_t e0, e1;
    JOTS_CUDA(cudaEventCreate(&e0)); JOTS_CUDA(cudaEventCreate(&e1));
    // JOTS_FORM_TIME_DOT=1 times the variant that also forms (x, A x); =2 that variant carrying the side stream of the
    // conjugate-gradient loop (x_sol += alpha d_old, zeroing of the next output) -- what the loop launches when it can
    const char* ev = std::getenv("JOTS_FORM_TIME_DOT");
    const int mode = (ev && form != JOTS_FORM_RESIDUAL) ? std::atoi(ev) : 0;
    double* d_dot = mode >= 1 ? c->d_scal.p + 63 : nullptr;
    DArr<double> sx, y2;
    if (mode >= 2) {
        sx.alloc(c->n); y2.alloc(c->n);
        c->zero(sx.p); c->zero(y2.p); c->zero(y);
        JOTS_CUDA(cudaMemsetAsync(c->d_scal.p + 62, 0, sizeof(double), c->stream));     // alpha = 0
    }
    double* Y[2] = {y, mode >= 2 ? y2.p : y};
    auto one = [&](int i) {
        if (mode >= 2) { const CgSide side{sx.p, x, c->d_scal.p + 62, Y[(i + 1) & 1], 0}; run_form(op, form, x, state, Y[i & 1], false, d_dot, nullptr, &side); }
        else run_form(op, form, x, state, y, false, d_dot);
    };
    one(0);
    c->sync();
    JOTS_CUDA(cudaEventRecord(e0, c->stream));
    for (int i = 0; i < reps; ++i) one(i + 1);
    JOTS_CUDA(cudaEventRecord(e1, c->stream));
    JOTS_CUDA(cudaEventSynchronize(e1));
    float t = 0;
    JOTS_CUDA(cudaEventElapsedTime(&t, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (ms) *ms = (double)t / (reps > 0 ? reps : 1);
    return 0;
    JOTS_CATCH_INT
}

// ---- partition / communicator -----------------------------------------------------------------
jots_part_t jots_partition_create(jots_mesh_t mesh, jots_space_t global, int nranks, int rank, const int32_t* dims) {
    JOTS_TRY
    std::vector<int> er;
    int d[3] = {0, 0, 0};
    if (dims) { d[0] = dims[0]; d[1] = dims[1]; d[2] = dims[2]; }
    partition_elements(mesh->m, nranks, dims ? d : nullptr, er);
    auto* h = new jots_part_s;
    h->p = make_partition(global->s, er, nranks, rank);
    return h;
    JOTS_CATCH_PTR
}
jots_part_t jots_partition_create_from_tables(int dim, int order, int64_t n_ldof, int64_t n_elem, const int32_t* gather, const double* elem_corners,
                                              int64_t n_bdr, const int32_t* bdr_gather, const double* bdr_corners, const int32_t* bdr_attr,
                                              const int64_t* ldof_global_id, int64_t n_true, const int32_t* ldof_of_true,
                                              int n_nbr, const int32_t* nbr_rank, const int64_t* nbr_ptr, const int32_t* nbr_ldofs,
                                              int n_bdr_attributes, const int32_t* bdr_attributes, int rank, int nranks) {
    JOTS_TRY
    if (!gather || !elem_corners || !ldof_global_id || !ldof_of_true) throw std::runtime_error("jots_partition_create_from_tables: null table");
    auto* h = new jots_part_s;
    h->p = make_partition_from_tables(dim, order, n_ldof, n_elem, gather, elem_corners, n_bdr, bdr_gather, bdr_corners, bdr_attr,
                                      ldof_global_id, n_true, ldof_of_true, n_nbr, nbr_rank, nbr_ptr, nbr_ldofs, n_bdr_attributes, bdr_attributes, rank, nranks);
    return h;
    JOTS_CATCH_PTR
}
void jots_partition_destroy(jots_part_t p) { delete p; }
int jots_partition_info(jots_part_t p, int64_t* n_local, int64_t* n_owned, int64_t* n_global, int64_t* n_elem, int* n_nbr, int64_t* n_shared) {
    if (n_local) *n_local = p->p.n_local;
    if (n_owned) *n_owned = p->p.n_owned;
    if (n_global) *n_global = p->p.n_global;
    if (n_elem) *n_elem = (int64_t)p->p.local_elems.size();
    if (n_nbr) *n_nbr = (int)p->p.nbr_rank.size();
    if (n_shared) *n_shared = (int64_t)p->p.red_dof.size();
    return 0;
}
jots_space_t jots_partition_space(jots_part_t p) {
    JOTS_TRY auto* h = new jots_space_s; h->s = p->p.local; return h; JOTS_CATCH_PTR
}
int jots_partition_copy(jots_part_t p, int64_t* global_id, int64_t* local_elems) {
    if (global_id) std::memcpy(global_id, p->p.global_id.data(), p->p.global_id.size() * sizeof(int64_t));
    if (local_elems) std::memcpy(local_elems, p->p.local_elems.data(), p->p.local_elems.size() * sizeof(int64_t));
    return 0;
}
int64_t jots_partition_neighbor(jots_part_t p, int i, int* rank, int32_t* local_idx) {
    if (i < 0 || i >= (int)p->p.nbr_rank.size()) { g_err = "neighbour index out of range"; return -1; }
    if (rank) *rank = p->p.nbr_rank[i];
    if (local_idx) std::memcpy(local_idx, p->p.nbr_dofs[i].data(), p->p.nbr_dofs[i].size() * sizeof(int));
    return (int64_t)p->p.nbr_dofs[i].size();
}
int64_t jots_partition_reduction(jots_part_t p, int32_t* red_dof, int32_t* red_ptr, int32_t* red_src) {
    const Partition& P = p->p;
    if (red_dof) std::memcpy(red_dof, P.red_dof.data(), P.red_dof.size() * sizeof(int));
    if (red_ptr) std::memcpy(red_ptr, P.red_ptr.data(), P.red_ptr.size() * sizeof(int));
    if (red_src) std::memcpy(red_src, P.red_src.data(), P.red_src.size() * sizeof(int));
    return (int64_t)P.red_src.size();
}
int jots_comm_unique_id(char out[128]) { JOTS_TRY comm_unique_id(out); return 0; JOTS_CATCH_INT }
jots_comm_t jots_comm_create(const char id[128], int rank, int nranks, int device) {
    JOTS_TRY auto* h = new jots_comm_s; h->c = comm_create(id, rank, nranks, device); return h; JOTS_CATCH_PTR
}
void jots_comm_destroy(jots_comm_t c) { if (c) { comm_destroy(c->c); delete c; } }
jots_ctx_t jots_ctx_create_part(jots_part_t p, int device, jots_comm_t comm) {
    JOTS_TRY
    auto* h = new jots_ctx_s;
    h->own.reset(new Ctx(p->p, device, comm ? comm->c : nullptr));
    h->c = h->own.get();
    return h;
    JOTS_CATCH_PTR
}

// ---- driver -----------------------------------------------------------------------------------
static void bind_views(jots_driver_s* h) {
    h->ctx_view.c = h->d->ctx.get();
    h->op_view.c = h->d->ctx.get();
    h->op_view.it = h->d->it.get();
    const std::string& st = h->d->cfg.sim_type;
    h->op_view.sim_type = st == "Linearized_Unsteady" ? JOTS_LINEARIZED_UNSTEADY : st == "Nonlinear_Unsteady" ? JOTS_NONLINEAR_UNSTEADY : JOTS_STEADY;
}
jots_driver_t jots_driver_create(const char* ini, int device) {
    JOTS_TRY
    Config cfg(ini);
    auto* h = new jots_driver_s;
    h->d.reset(new JOTSDriver(cfg, device));
    bind_views(h);
    return h;
    JOTS_CATCH_PTR
}
jots_driver_t jots_driver_create_on_mesh(const char* ini, jots_mesh_t mesh, int device) {
    JOTS_TRY
    Config cfg(ini);
    auto* h = new jots_driver_s;
    h->d.reset(new JOTSDriver(cfg, device, &mesh->m));
    bind_views(h);
    return h;
    JOTS_CATCH_PTR
}
jots_driver_t jots_driver_create_dist(const char* ini, jots_mesh_t mesh, int device, jots_comm_t comm) {
    JOTS_TRY
    Config cfg(ini);
    auto* h = new jots_driver_s;
    h->d.reset(new JOTSDriver(cfg, device, mesh ? &mesh->m : nullptr, comm ? comm->c : nullptr));
    bind_views(h);
    return h;
    JOTS_CATCH_PTR
}
int jots_driver_copy_global_ids(jots_driver_t d, int64_t* out, int64_t* n_owned) {
    JOTS_TRY
    Ctx* c = d->d->ctx.get();
    if (d->d->part) std::memcpy(out, d->d->part->global_id.data(), c->n * sizeof(int64_t));
    else for (int64_t i = 0; i < c->n; ++i) out[i] = i;
    if (n_owned) *n_owned = c->n_owned;
    return 0;
    JOTS_CATCH_INT
}
void jots_driver_destroy(jots_driver_t d) { delete d; }
int jots_driver_run(jots_driver_t d, int max_steps, int* done) {
    JOTS_TRY
    int n = d->d->Run(max_steps);
    if (done) *done = n;
    return 0;
    JOTS_CATCH_INT
}
int64_t jots_driver_size(jots_driver_t d) { return d->d->ctx->n; }
int jots_driver_get_u(jots_driver_t d, double* u, int loc) {
    JOTS_TRY
    Ctx* c = d->d->ctx.get();
    JOTS_CUDA(cudaMemcpyAsync(u, d->d->u.p, c->n * sizeof(double), loc == JOTS_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    c->sync();
    return 0;
    JOTS_CATCH_INT
}
int jots_driver_set_u(jots_driver_t d, const double* u, int loc) {
    JOTS_TRY
    Ctx* c = d->d->ctx.get();
    JOTS_CUDA(cudaMemcpyAsync(d->d->u.p, u, c->n * sizeof(double), loc == JOTS_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    c->sync();
    return 0;
    JOTS_CATCH_INT
}
double jots_driver_time(jots_driver_t d) { return d->d->time; }
jots_ctx_t jots_driver_ctx(jots_driver_t d) { return &d->ctx_view; }
jots_op_t jots_driver_op(jots_driver_t d) { return &d->op_view; }
int jots_driver_step_count(jots_driver_t d) { return d->d->it_num; }
jots_space_t jots_driver_space(jots_driver_t d) {
    JOTS_TRY auto* h = new jots_space_s; h->s = d->d->ctx->space; return h; JOTS_CATCH_PTR
}

}  // extern "C"
