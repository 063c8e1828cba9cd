#include "ctx.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include <nvtx3/nvToolsExt.h>

#include "basis.hpp"
#include "halo.hpp"
#include "launch.hpp"
#include "partition.hpp"

namespace jots {

void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
    if (e == cudaSuccess) return;
    throw CudaError(std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " (" + file + ":" + std::to_string(line) + ")");
}

NvtxRange::NvtxRange(const char* name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }

void BCSpec::update(double t) {
    // SinusoidalBC::UpdateCoeff: A sin(w t + phase) + shift (boundary_condition.cpp:65-75)
    if (constant() || precice()) value = v[0];      // PreciceBC: the default value until the adapter writes read data
    else value = v[0] * std::sin(v[1] * t + v[2]) + v[3];
}

static void fill_tables(TablesRT& t, int dim, int p) {
    std::memset(&t, 0, sizeof(t));
    const int D = p + 1, Q = domain_quad_points(dim, p);
    t.D = D; t.Q = Q; t.DZ = dim == 3 ? D : 1; t.QZ = dim == 3 ? Q : 1;
    std::vector<double> nodes = gauss_lobatto_01(D), qx, qw, B, G, Bq, Gq;
    gauss_legendre_01(Q, qx, qw);
    lagrange_tab(nodes, qx, B, G);
    lagrange_tab(qx, qx, Bq, Gq);
    for (int i = 0; i < Q * D; ++i) { t.B[i] = B[i]; t.G[i] = G[i]; }
    for (int i = 0; i < Q * Q; ++i) t.Gq[i] = Gq[i];
    for (int i = 0; i < Q; ++i) t.W[i] = qw[i];
    if (dim == 3) {
        for (int i = 0; i < Q * D; ++i) { t.Bz[i] = B[i]; t.Gz[i] = G[i]; }
        for (int i = 0; i < Q * Q; ++i) t.Gqz[i] = Gq[i];
        for (int i = 0; i < Q; ++i) t.Wz[i] = qw[i];
    } else {
        t.Bz[0] = 1.0; t.Gz[0] = 0.0; t.Gqz[0] = 0.0; t.Wz[0] = 1.0;
    }
}

Ctx::Ctx(const Space& s, int dev) : device(dev), space(s) { init(s, dev); }

Ctx::Ctx(const Partition& p, int dev, Comm* comm) : device(dev), space(p.local) {
    init(p.local, dev);
    n_owned = p.n_owned;
    if (comm) halo = halo_create(p, comm);
}

void Ctx::init(const Space& s, int dev) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw CudaError("no CUDA device available: the conduction solve has no CPU fallback");
    JOTS_CUDA(cudaSetDevice(dev));
    JOTS_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    own_stream = true;
    n = s.ndof; n_owned = s.ndof;
    if (const char* ev = std::getenv("JOTS_DISABLE_SLAB")) use_slab = !(ev[0] == '1');
    if (const char* ev = std::getenv("JOTS_SLAB_VARIANT")) slab_variant = std::atoi(ev);
    fill_tables(tabs, s.dim, s.p);
    // geometry: affine elements get one record each (one shared record when all are identical);
    // general multilinear meshes get a record per quadrature point (d_geomq)
    const int dim = s.dim, nv = 1 << dim;
    std::vector<double> geom((size_t)s.nelem * 10);
    bool uniform = true, affine = true;
    geom_diag = true;
    auto record = [](const double J[3][3], double* g) {
        const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                           J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
        if (!(det > 0)) throw std::runtime_error("non-positive element Jacobian");
        // invJ[a][i] = d xi_a / d x_i
        g[0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det; g[1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det; g[2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
        g[3] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det; g[4] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det; g[5] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
        g[6] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det; g[7] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det; g[8] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
        g[9] = det;
    };
    for (int64_t el = 0; el < s.nelem; ++el) {
        const double* X = &s.elem_corner[(size_t)el * nv * dim];
        double J[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int a = 0; a < dim; ++a)
            for (int i = 0; i < dim; ++i) J[i][a] = X[(1 << a) * dim + i] - X[i];
        double scale = 0.0;
        for (int a = 0; a < dim; ++a) for (int i = 0; i < dim; ++i) scale = std::max(scale, std::fabs(J[i][a]));
        for (int c = 0; c < nv && affine; ++c)
            for (int i = 0; i < dim; ++i) {
                double x = X[i];
                for (int a = 0; a < dim; ++a) if ((c >> a) & 1) x += J[i][a];
                if (std::fabs(x - X[c * dim + i]) > 1e-10 * scale) { affine = false; break; }
            }
        double* g = &geom[(size_t)el * 10];
        record(J, g);
        {
            // metric G = invJ invJ^T diagonal?
            const double g01 = g[0] * g[3] + g[1] * g[4] + g[2] * g[5], g02 = g[0] * g[6] + g[1] * g[7] + g[2] * g[8],
                         g12 = g[3] * g[6] + g[4] * g[7] + g[5] * g[8];
            const double gs = g[0] * g[0] + g[1] * g[1] + g[2] * g[2] + g[3] * g[3] + g[4] * g[4] + g[5] * g[5] + g[6] * g[6] + g[7] * g[7] + g[8] * g[8];
            if (std::fabs(g01) + std::fabs(g02) + std::fabs(g12) > 1e-13 * gs) geom_diag = false;
        }
        if (uniform && el > 0)
            for (int i = 0; i < 10; ++i)
                if (std::fabs(g[i] - geom[i]) > 1e-12 * std::max(std::fabs(geom[i]), std::fabs(geom[i < 9 ? 4 * (i / 3) : 9]))) { uniform = false; break; }
    }
    if (!affine) {
        // Jacobian of the multilinear map at every quadrature point (the oracle / MFEM evaluate the
        // same map: ElementTransformation of a d-linear element)
        uniform = false; geom_diag = false;
        const int Q = tabs.Q, QZ = tabs.QZ, NQ = Q * Q * QZ;
        std::vector<double> qx, qw;
        gauss_legendre_01(Q, qx, qw);
        std::vector<double> gq((size_t)s.nelem * NQ * 10);
        for (int64_t el = 0; el < s.nelem; ++el) {
            const double* X = &s.elem_corner[(size_t)el * nv * dim];
            for (int q = 0; q < NQ; ++q) {
                const double xi[3] = {qx[q % Q], qx[(q / Q) % Q], dim == 3 ? qx[q / (Q * Q)] : 0.0};
                double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 1}};
                if (dim == 2) { J[2][2] = 1.0; }
                else J[2][2] = 0.0;
                for (int c = 0; c < nv; ++c)
                    for (int a = 0; a < dim; ++a) {
                        double dn = ((c >> a) & 1) ? 1.0 : -1.0;
                        for (int b = 0; b < dim; ++b) if (b != a) dn *= ((c >> b) & 1) ? xi[b] : 1.0 - xi[b];
                        for (int i = 0; i < dim; ++i) J[i][a] += dn * X[c * dim + i];
                    }
                record(J, &gq[((size_t)el * NQ + q) * 10]);
            }
        }
        d_geomq.upload(gq);
    }
    if (uniform && s.nelem > 0) { geom.resize(10); geom_stride = 0; }
    d_geom.upload(geom);
    d_gather.upload(s.gather);
    d_gather_ess.upload(s.gather);
    d_bdr_gather.upload(s.bdr_gather);
    d_bdr_attr.upload(s.bdr_attr);
    d_bdr_corner.upload(s.bdr_corner);
    d_scal.alloc(64);
    d_partial.alloc((size_t)DOT_MAX_BLOCKS * 4);
    d_counter.alloc(1);
    JOTS_CUDA(cudaMemset(d_counter.p, 0, sizeof(unsigned int)));
    d_sched.alloc(2);
    JOTS_CUDA(cudaMemset(d_sched.p, 0, 2 * sizeof(unsigned int)));
    JOTS_CUDA(cudaMallocHost((void**)&h_scal, 64 * sizeof(double)));
    std::vector<unsigned char> em((size_t)n, 0);
    d_essmark.upload(em);
    Poly one; one.n = 1; std::memset(one.c, 0, sizeof(one.c)); one.c[0] = 1.0;
    mat.rho = one; mat.C = one; mat.k = one;
}

Ctx::~Ctx() {
    cudaSetDevice(device);
    if (h_scal) cudaFreeHost(h_scal);
    if (own_stream && stream) cudaStreamDestroy(stream);
    halo_destroy(halo);
}

void Ctx::set_stream(cudaStream_t s) {
    JOTS_CUDA(cudaSetDevice(device));
    sync();
    if (own_stream && stream) cudaStreamDestroy(stream);
    own_stream = false;
    stream = s;
}

void Ctx::expand_owned(double* v) {
    if (!halo) return;
    // collective: a rank that owns every DOF it touches (the lowest rank) still has to serve its neighbours
    if (n_owned < n) JOTS_CUDA(cudaMemsetAsync(v + n_owned, 0, (size_t)(n - n_owned) * sizeof(double), stream));
    halo_sum_exchange(*this, v);
}

void Ctx::set_materials(const Poly& rho, const Poly& C, const Poly& k) {
    mat.rho = rho; mat.C = C; mat.k = k;
}

void Ctx::set_bcs(const std::vector<BCSpec>& b) {
    bcs = b;
    // dbc_bdr[i] = 1 for essential BC i; marker index i <-> attribute i+1 (jots_iterator.cpp:12-30)
    std::vector<int> attrs;
    bc_dofs.assign(bcs.size(), {});
    d_bc_dofs.clear();
    d_bc_dofs.resize(bcs.size());
    for (size_t i = 0; i < bcs.size(); ++i) {
        std::vector<int> one{(int)i + 1};
        space.essential_dofs(one, bc_dofs[i]);
        if (space.sync_bdr_markers && halo) {
            // caller-supplied subdomain: a shared DOF may lie on a boundary face that only a neighbour's element has; sum the
            // 0/1 markers over the interface like ParFiniteElementSpace::GetEssentialTrueDofs synchronises its marker
            std::vector<double> mk((size_t)n, 0.0);
            for (int d : bc_dofs[i]) mk[d] = 1.0;
            DArr<double> dm;
            dm.upload(mk);
            halo_sum_exchange(*this, dm.p);
            JOTS_CUDA(cudaMemcpyAsync(mk.data(), dm.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, stream));
            sync();
            bc_dofs[i].clear();
            for (int64_t d = 0; d < n; ++d) if (mk[d] != 0.0) bc_dofs[i].push_back((int)d);
        }
        d_bc_dofs[i].upload(bc_dofs[i]);
        if (bcs[i].essential()) attrs.push_back((int)i + 1);
    }
    if (space.sync_bdr_markers && halo) {
        ess.clear();
        for (int a : attrs) ess.insert(ess.end(), bc_dofs[a - 1].begin(), bc_dofs[a - 1].end());
        std::sort(ess.begin(), ess.end());
        ess.erase(std::unique(ess.begin(), ess.end()), ess.end());
    } else space.essential_dofs(attrs, ess);
    d_ess.upload(ess);
    {
        std::vector<int> eo;
        for (int d : ess) if (d < n_owned) eo.push_back(d);
        n_ess_owned = (int)eo.size();
        d_ess_owned.upload(eo);
    }
    std::vector<unsigned char> em((size_t)n, 0);
    for (int d : ess) em[d] = 1;
    d_essmark.upload(em);
    std::vector<int> ge(space.gather);
    for (auto& g : ge) if (em[g]) g = ~g;
    d_gather_ess.upload(ge);
    d_gattr.alloc(std::max<size_t>(bcs.size(), 1));
    // preCICE BCs: coefficient field, boundary-element node lists
    bc_field.clear(); bc_field.resize(bcs.size());
    bc_faces.assign(bcs.size(), {}); bc_node_dofs.assign(bcs.size(), {});
    d_bc_faces.clear(); d_bc_faces.resize(bcs.size());
    d_bc_node_dofs.clear(); d_bc_node_dofs.resize(bcs.size());
    std::vector<const double*> gn(std::max<size_t>(bcs.size(), 1), nullptr);
    bool any_precice = false;
    for (size_t i = 0; i < bcs.size(); ++i) {
        if (!bcs[i].precice()) continue;
        any_precice = true;
        bc_field[i].alloc(n);
        v_fill(stream, n, bcs[i].v[0], bc_field[i].p);
        for (int64_t b = 0; b < space.nbdr; ++b) {
            if (space.bdr_attr[b] != (int)i + 1) continue;
            bc_faces[i].push_back((int)b);
            for (int l = 0; l < space.nfl; ++l) bc_node_dofs[i].push_back(space.bdr_gather[b * space.nfl + l]);
        }
        d_bc_faces[i].upload(bc_faces[i]);
        d_bc_node_dofs[i].upload(bc_node_dofs[i]);
        if (!bcs[i].essential()) gn[i] = bc_field[i].p;
    }
    d_gnodal.upload(gn);
    if (any_precice && d_bdr_elem.n != (size_t)space.nbdr) {
        fill_boundary_adjacency(space);        // spaces from caller-supplied tables: matched through the DOF ids
        std::vector<int> nl;
        boundary_node_locals(space, nl);
        d_bdr_node_loc.upload(nl);
        std::vector<int> be(space.nbdr, -1), bf(space.nbdr, -1);
        for (int64_t b = 0; b < space.nbdr && b < (int64_t)space.bdr_elem.size(); ++b) { be[b] = (int)space.bdr_elem[b]; bf[b] = space.bdr_face[b]; }
        d_bdr_elem.upload(be); d_bdr_face.upload(bf);
    }
    sync();
    build_patches();
}

// Element patches of apply_patch.cuh (make_patches, patches.cpp): rebuilt with the essential set
void Ctx::build_patches() {
    d_patches.release();
    n_patches = 0; patch_fill = 0.0;
    if (space.dim != 3 || !geom_diag || !use_slab || slab_variant != 7 || space.nelem == 0) return;
    int px = 0, py = 0;
    patch_shape(space.p, px, py);
    if (px <= 0) return;
    std::vector<int> tab;
    double fill = 0.0;
    const int64_t np = make_patches(space, ess, px, py, tab, &fill);
    double min_fill = 0.6;
    if (const char* ev = std::getenv("JOTS_PATCH_MIN_FILL")) min_fill = std::atof(ev);
    if (np <= 0 || np > 0x7fffffffLL || fill < min_fill) return;
    d_patches.upload(tab);
    n_patches = (int)np;
    patch_fill = fill;
}

bool Ctx::any_flux() const {
    for (auto& b : bcs) if (!b.essential() && (b.value != 0.0 || b.precice())) return true;
    return false;
}

void Ctx::upload_flux_values() {
    std::vector<double> g(std::max<size_t>(bcs.size(), 1), 0.0);
    for (size_t i = 0; i < bcs.size(); ++i) g[i] = (bcs[i].essential() || bcs[i].precice()) ? 0.0 : bcs[i].value;   // preCICE: nodal field
    JOTS_CUDA(cudaMemcpyAsync(d_gattr.p, g.data(), g.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
    JOTS_CUDA(cudaStreamSynchronize(stream));
}

FormArgs Ctx::make_args(const FormSpec& f, const double* x, double* y, bool for_diag) const {
    FormArgs a;
    std::memset(&a, 0, sizeof(a));
    a.gather = d_gather_ess.p;
    a.geom = d_geom.p; a.geom_stride = geom_stride;
    a.geom_diag = (geom_diag && use_slab) ? 1 : 0;
    a.geomq = d_geomq.p;
    a.x = x; a.z = f.state; a.y = y;
    a.zc = (f.coef_state && f.coef_state != f.state) ? f.coef_state : nullptr;
    a.nelem = (int)space.nelem;
    a.mask_in = f.mask_in; a.mask_out = f.mask_out;
    a.diag_one = (f.diag_one && f.mask_out && !halo && !for_diag) ? 1 : 0;   // multi-GPU: after the interface sum
    a.cm = f.cm; a.ck = f.ck; a.cb = f.cb; a.m0 = f.m0; a.a0 = f.a0; a.inv_rc = f.inv_rc;
    a.mat = mat;
    a.nvec = n;
    a.patches = (n_patches > 0 && !for_diag) ? d_patches.p : nullptr;
    a.n_patches = n_patches;
    a.skip = for_diag ? nullptr : apply_skip;
    a.sched = d_sched.p;
    return a;
}

static void fill_face_tables(FaceArgs& fa, int p, int nq) {
    const int D = p + 1;
    std::vector<double> nodes = gauss_lobatto_01(D), qx, qw, B, G;
    gauss_legendre_01(nq, qx, qw);
    lagrange_tab(nodes, qx, B, G);
    for (int i = 0; i < nq * D; ++i) fa.Bf[i] = B[i];
    for (int i = 0; i < nq; ++i) { fa.Wf[i] = qw[i]; fa.xq[i] = qx[i]; }
    fa.nq = nq;
}

static FaceArgs base_face_args(const Ctx& c, int nq) {
    FaceArgs fa;
    std::memset(&fa, 0, sizeof(fa));
    fa.dim = c.space.dim; fa.D = c.space.D; fa.nfl = c.space.nfl; fa.nbdr = (int)c.space.nbdr;
    fa.bdr_gather = c.d_bdr_gather.p; fa.bdr_corner = c.d_bdr_corner.p; fa.bdr_attr = c.d_bdr_attr.p;
    fa.g_attr = c.d_gattr.p; fa.n_attr = (int)c.bcs.size();
    fa.g_nodal = c.d_gnodal.p;
    fa.ess = c.d_essmark.p;
    fa.mat = c.mat;
    fa.scale = 1.0; fa.div_const = 1.0;
    fill_face_tables(fa, c.space.p, nq);
    return fa;
}

// Fused (x, A x): symmetric-elimination forms whose apply is ONE element kernel that honours FormArgs::dot -- the two-phase
// patch kernel on Q2 lattices, the gather-table kernel where it has two staging buffers (Q1, Q3, Q2 with rho(T) / C(T) on
// meshes without patches).  Returns the number of tiles of that launch (0: not fused) and the entries per tile its side
// stream can carry.
static int form_dot_tiles(const Ctx& c, const FormSpec& f, const FormArgs& a, int* side_capacity) {
    const bool face_terms = f.bdr_mass_scale != 0.0 && c.any_flux();
    if (face_terms || !a.geom_diag || c.space.dim != 3 || a.zc || f.mask_in != f.mask_out) return 0;
    if (a.patches && a.n_patches > 0 && c.space.p == 2) {      // apply_form takes the patch kernels
        if (!patch_kernel_fuses_dot(c.space.p, f.op, f.cf, a)) return 0;
        if (side_capacity) *side_capacity = patch_kernel_side_capacity(c.space.p, f.op, f.cf, a);
        return a.n_patches;
    }
    return slab_kernel_dot_tiles(c.space.p, f.op, f.cf, a, side_capacity);
}

bool Ctx::ess_rows_are_zero(const double* v) {
    if (n_ess_owned == 0) return true;
    double* acc = d_scal.p + 61;
    JOTS_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), stream));
    v_zero_essdot(stream, 0, nullptr, n_ess_owned, d_ess_owned.p, v, acc, nullptr);     // sum of squares over the listed entries
    ++stats.kernel_launches;
    double h = 1.0;
    JOTS_CUDA(cudaMemcpyAsync(&h, acc, sizeof(double), cudaMemcpyDeviceToHost, stream));
    sync();
    return h == 0.0;
}

int Ctx::form_side_chunk(const FormSpec& f) const {
    const FormArgs a = make_args(f, nullptr, nullptr, false);
    int cap = 0;
    const int tiles = form_dot_tiles(*this, f, a, &cap);
    if (tiles <= 0 || n > 2000000000LL) return 0;      // the kernels index the slices with 32-bit arithmetic
    int64_t chunk = (n + tiles - 1) / tiles;
    chunk += chunk & 1;
    return chunk <= cap ? (int)chunk : 0;
}

bool Ctx::apply_form(const FormSpec& f, const double* x, double* y, double* dot, const CgSide* side) {
    FormArgs a = make_args(f, x, y, false);
    const bool face_terms = f.bdr_mass_scale != 0.0 && any_flux();
    const bool fuse = dot && form_dot_tiles(*this, f, a, nullptr) > 0;
    if (side) {
        a.side_chunk = fuse ? form_side_chunk(f) : 0;
        auto misaligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; };
        if (!a.side_chunk || misaligned(side->x) || misaligned(side->d) || misaligned(side->zero))
            throw NoSideStream("apply_form: this form / these vectors cannot carry the CG side stream");
        a.side_x = side->x; a.side_d = side->d; a.side_alpha = side->alpha; a.side_zero = side->zero;
    }
    if (fuse) {
        a.dot = dot;
        // with a side stream y was zeroed by the previous launch: only the unit-diagonal rows' share of (x, A x) is left
        const int n_idx = (f.diag_one && f.mask_out) ? n_ess_owned : 0;
        if (!side) v_zero_essdot(stream, n, y, n_idx, d_ess_owned.p, x, dot, a.skip);
        else if (n_idx && !side->ess_zero) v_zero_essdot(stream, 0, y, n_idx, d_ess_owned.p, x, dot, a.skip);
    } else zero(y);
    // lattice meshes: unique-pencil patch kernels; otherwise / when not instantiated the gather-table kernels
    int rc = 1;
    if (a.patches && a.geom_diag && space.dim == 3 && !a.zc) {
        rc = launch_form_apply_patch(space.p, f.op, f.cf, a, tabs, stream);
        if (!rc) ++stats.patch_applies;
    }
    if (rc) rc = launch_form_apply(space.dim, space.p, f.op, f.cf, a, tabs, stream);
    if (rc) throw std::runtime_error("form (op=" + std::to_string(f.op) + ", cf=" + std::to_string(f.cf) + ") is not instantiated, rc=" + std::to_string(rc));
    stats.kernel_launches += (side && (side->ess_zero || !(f.diag_one && f.mask_out && n_ess_owned))) ? 1 : 2;
    last_kernel = last_element_kernel();
    if (fuse && !std::strstr(last_kernel, ",dot")) throw std::runtime_error(std::string("apply_form: (x, A x) was promised by a kernel that cannot form it: ") + last_kernel);
    if (face_terms) {
        FaceArgs fa = base_face_args(*this, bdr_mass_quad_points(space.p));
        fa.z = f.state; fa.x = x; fa.y = y; fa.mask_in = f.mask_in; fa.mask_out = f.mask_out; fa.scale = f.bdr_mass_scale;
        launch_face(FACE_MASS_APPLY, fa, stream);
        ++stats.kernel_launches;
    }
    if (halo) halo_sum_exchange(*this, y);
    if (f.diag_one && !a.diag_one) { v_copy_idx(stream, (int)ess.size(), d_ess.p, x, y); stats.kernel_launches += !ess.empty(); }
    JOTS_CUDA(cudaGetLastError());
    ++stats.applies;
    if (fuse) ++stats.fused_dots;
    if (side) ++stats.side_applies;
    return fuse;
}

void Ctx::diag_form(const FormSpec& f, double* d) {
    NvtxRange range("jots:diagonal");
    zero(d);
    FormArgs a = make_args(f, nullptr, d, true);
    a.mask_out = (f.mask_out || f.diag_one) ? 1 : 0;
    int rc = launch_form_diag(space.dim, space.p, f.op == OP_RES ? OP_LIN : f.op, f.cf, a, tabs, stream);
    if (rc) throw std::runtime_error("diagonal of form (op=" + std::to_string(f.op) + ", cf=" + std::to_string(f.cf) + ") is not instantiated");
    stats.kernel_launches += 2;
    if (f.bdr_mass_scale != 0.0 && any_flux()) {
        FaceArgs fa = base_face_args(*this, bdr_mass_quad_points(space.p));
        fa.z = f.state; fa.y = d; fa.mask_out = a.mask_out; fa.scale = f.bdr_mass_scale;
        launch_face(FACE_MASS_DIAG, fa, stream);
        ++stats.kernel_launches;
    }
    if (halo) halo_sum_exchange(*this, d);
    if (f.diag_one) { v_set_idx(stream, (int)ess.size(), d_ess.p, 1.0, d); stats.kernel_launches += !ess.empty(); }
    JOTS_CUDA(cudaGetLastError());
    ++stats.diag_assemblies;
}

void Ctx::neumann_vector(double div_const, const double* state, double* b, bool zero_first, double scale, bool mask_out) {
    if (zero_first) zero(b);
    FaceArgs fa = base_face_args(*this, bdr_lf_quad_points(space.p));
    fa.z = state; fa.y = b; fa.div_const = div_const; fa.scale = scale; fa.mask_out = mask_out;
    launch_face(FACE_LOAD, fa, stream);
    ++stats.kernel_launches;
    if (halo) {
        if (!zero_first) throw std::runtime_error("accumulating Neumann assembly is not available on partitioned contexts");
        halo_sum_exchange(*this, b);
    }
    JOTS_CUDA(cudaGetLastError());
}

void Ctx::fetch_scalars(int nv, double* out) {
    if (halo) halo_allreduce(*this, d_scal.p, nv);
    JOTS_CUDA(cudaMemcpyAsync(h_scal, d_scal.p, nv * sizeof(double), cudaMemcpyDeviceToHost, stream));
    JOTS_CUDA(cudaStreamSynchronize(stream));
    for (int i = 0; i < nv; ++i) out[i] = h_scal[i];
}

void Ctx::allreduce_scalars(double* d_vals, int nv) {
    if (halo) halo_allreduce(*this, d_vals, nv);
}

void Ctx::copy_scalars(int nv, double* out) {
    JOTS_CUDA(cudaMemcpyAsync(h_scal, d_scal.p, nv * sizeof(double), cudaMemcpyDeviceToHost, stream));
    JOTS_CUDA(cudaStreamSynchronize(stream));
    for (int i = 0; i < nv; ++i) out[i] = h_scal[i];
}

void Ctx::dots(int nv, const double* const* a, const double* const* b, double* out) {
    v_dots(stream, n_owned, nv, a, b, dot_scratch(), d_scal.p);
    ++stats.kernel_launches; stats.dots += nv;
    fetch_scalars(nv, out);
}

double Ctx::dot(const double* a, const double* b) {
    double r;
    const double* aa[1] = {a};
    const double* bb[1] = {b};
    dots(1, aa, bb, &r);
    return r;
}

double Ctx::norm(const double* a) { return std::sqrt(dot(a, a)); }

double Ctx::max_local(const double* a) {
    v_max(stream, n, a, dot_scratch(), d_scal.p);
    ++stats.kernel_launches;
    JOTS_CUDA(cudaMemcpyAsync(h_scal, d_scal.p, sizeof(double), cudaMemcpyDeviceToHost, stream));
    JOTS_CUDA(cudaStreamSynchronize(stream));
    return h_scal[0];
}

void Ctx::apply_dirichlet(int bc, double* u) {
    // ProjectBdrCoefficient(coeff, marker_bc) (jots_driver.cpp:692): a uniform value, or the nodal field of a preCICE BC
    const int cnt = (int)bc_dofs[bc].size();
    if (bcs[bc].precice()) v_copy_idx(stream, cnt, d_bc_dofs[bc].p, bc_field[bc].p, u);
    else v_set_idx(stream, cnt, d_bc_dofs[bc].p, bcs[bc].value, u);
    stats.kernel_launches += cnt > 0;
}

void Ctx::precice_bc_coords(int bc, double* xyz) const {
    if (bc < 0 || bc >= (int)bcs.size() || !bcs[bc].precice()) throw std::runtime_error("not a preCICE boundary condition");
    std::vector<double> X;
    space.node_coordinates(X);
    const std::vector<int>& nd = bc_node_dofs[bc];
    for (size_t i = 0; i < nd.size(); ++i)
        for (int d = 0; d < space.dim; ++d) xyz[i * space.dim + d] = X[(size_t)nd[i] * space.dim + d];
}

void Ctx::precice_bc_set_read_data(int bc, const double* vals) {
    if (bc < 0 || bc >= (int)bcs.size() || !bcs[bc].precice()) throw std::runtime_error("not a preCICE boundary condition");
    // coeff_gf.SetSubVector(bdr_dof_indices, read_data_arr): a DOF listed by several boundary elements keeps the last value
    const std::vector<int>& nd = bc_node_dofs[bc];
    std::vector<int> dofs;
    std::vector<double> v;
    {
        std::vector<std::pair<int, double>> last(nd.size());
        for (size_t i = 0; i < nd.size(); ++i) last[i] = {nd[i], vals[i]};
        std::stable_sort(last.begin(), last.end(), [](const std::pair<int, double>& a, const std::pair<int, double>& b) { return a.first < b.first; });
        for (size_t i = 0; i < last.size(); ++i)
            if (i + 1 == last.size() || last[i + 1].first != last[i].first) { dofs.push_back(last[i].first); v.push_back(last[i].second); }
    }
    DArr<int> dd; DArr<double> dv;
    dd.upload(dofs); dv.upload(v);
    v_scatter_idx(stream, (int)dofs.size(), dd.p, dv.p, bc_field[bc].p);
    ++stats.kernel_launches;
    sync();
}

void Ctx::precice_bc_write_data(int bc, const double* u, double* out) {
    if (bc < 0 || bc >= (int)bcs.size() || !bcs[bc].precice()) throw std::runtime_error("not a preCICE boundary condition");
    const size_t cnt = bc_node_dofs[bc].size();
    DArr<double> d_out;
    d_out.alloc(std::max<size_t>(cnt, 1));
    if (bcs[bc].essential()) {
        if (d_geomq.p) throw std::runtime_error("wall heat flux: general multilinear (non-affine) elements are not supported");
        for (int b : bc_faces[bc]) if (b >= (int)space.bdr_elem.size() || space.bdr_elem[b] < 0) throw std::runtime_error("wall heat flux: the space has no boundary-element adjacency (caller-supplied tables)");
        WallFluxArgs a;
        std::memset(&a, 0, sizeof(a));
        a.dim = space.dim; a.D = space.D; a.nloc = space.nloc; a.nfl = space.nfl; a.nfaces = (int)bc_faces[bc].size();
        a.faces = d_bc_faces[bc].p; a.bdr_elem = d_bdr_elem.p; a.bdr_face = d_bdr_face.p; a.node_loc = d_bdr_node_loc.p; a.gather = d_gather_ess.p;
        a.geom = d_geom.p; a.geom_stride = geom_stride; a.u = u; a.out = d_out.p; a.k = mat.k;
        std::vector<double> nodes = gauss_lobatto_01(space.D), Bn, Gn;
        lagrange_tab(nodes, nodes, Bn, Gn);
        for (int i = 0; i < space.D * space.D; ++i) a.Dn[i] = Gn[i];
        launch_wall_flux(a, stream);
    } else {
        v_pack(stream, (int)cnt, d_bc_node_dofs[bc].p, u, d_out.p);     // PreciceHeatFluxBC: u_gf.GetSubVector(bdr_dof_indices)
    }
    ++stats.kernel_launches;
    JOTS_CUDA(cudaMemcpyAsync(out, d_out.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, stream));
    sync();
}

double Ctx::max_global(const double* a) {
    v_max(stream, n, a, dot_scratch(), d_scal.p);
    ++stats.kernel_launches;
    if (halo) halo_allreduce_max(*this, d_scal.p, 1);
    JOTS_CUDA(cudaMemcpyAsync(h_scal, d_scal.p, sizeof(double), cudaMemcpyDeviceToHost, stream));
    JOTS_CUDA(cudaStreamSynchronize(stream));
    return h_scal[0];
}

int Ctx::rank() const { return halo ? halo_rank(*this) : 0; }

}  // namespace jots
