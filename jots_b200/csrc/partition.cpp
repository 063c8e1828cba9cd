#include "partition.hpp"

#include "basis.hpp"

#include <algorithm>
#include <numeric>
#include <stdexcept>

namespace jots {

static void box_dims(int nranks, const int64_t bn[3], int dims[3]) {
    // factor nranks into px*py*pz, repeatedly splitting the direction with most elements per part
    dims[0] = dims[1] = dims[2] = 1;
    int r = nranks;
    for (int f = 2; r > 1;) {
        if (r % f) { ++f; continue; }
        int best = 0;
        double bestv = -1;
        for (int a = 0; a < 3; ++a) {
            const double v = (double)bn[a] / dims[a];
            if (v > bestv + 1e-12) { bestv = v; best = a; }
        }
        dims[best] *= f;
        r /= f;
    }
}

static void rcb(const std::vector<double>& cen, int dim, std::vector<int64_t>& ids, int64_t lo, int64_t hi, int r0, int nr,
                std::vector<int>& out) {
    if (nr == 1) { for (int64_t i = lo; i < hi; ++i) out[ids[i]] = r0; return; }
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int64_t i = lo; i < hi; ++i)
        for (int a = 0; a < dim; ++a) { mn[a] = std::min(mn[a], cen[ids[i] * dim + a]); mx[a] = std::max(mx[a], cen[ids[i] * dim + a]); }
    int ax = 0;
    for (int a = 1; a < dim; ++a) if (mx[a] - mn[a] > (mx[ax] - mn[ax]) * (1 + 1e-12)) ax = a;
    const int nl = nr / 2;
    const int64_t mid = lo + (hi - lo) * nl / nr;
    std::stable_sort(ids.begin() + lo, ids.begin() + hi, [&](int64_t x, int64_t y) { return cen[x * dim + ax] < cen[y * dim + ax]; });
    rcb(cen, dim, ids, lo, mid, r0, nl, out);
    rcb(cen, dim, ids, mid, hi, r0 + nl, nr - nl, out);
}

void partition_elements(const Mesh& m, int nranks, const int* dims_in, std::vector<int>& elem_rank) {
    const int64_t ne = m.n_elem();
    elem_rank.assign(ne, 0);
    if (nranks <= 1) return;
    if (m.is_brick) {
        int dims[3];
        if (dims_in && dims_in[0] > 0) { dims[0] = dims_in[0]; dims[1] = dims_in[1]; dims[2] = dims_in[2]; }
        else box_dims(nranks, m.bn, dims);
        if (dims[0] * dims[1] * dims[2] != nranks) throw std::runtime_error("partition dims do not multiply to the number of ranks");
        for (int a = 0; a < 3; ++a) if (dims[a] > m.bn[a]) throw std::runtime_error("more subdomains than elements in one direction");
        int64_t e = 0;
        for (int64_t k = 0; k < m.bn[2]; ++k) for (int64_t j = 0; j < m.bn[1]; ++j) for (int64_t i = 0; i < m.bn[0]; ++i, ++e) {
            const int pi = (int)(i * dims[0] / m.bn[0]), pj = (int)(j * dims[1] / m.bn[1]), pk = (int)(k * dims[2] / m.bn[2]);
            elem_rank[e] = pi + dims[0] * (pj + dims[1] * pk);
        }
        return;
    }
    const int dim = m.dim, nv = m.nvpe();
    std::vector<double> cen(ne * dim, 0.0);
    for (int64_t e = 0; e < ne; ++e)
        for (int v = 0; v < nv; ++v)
            for (int a = 0; a < dim; ++a) cen[e * dim + a] += m.verts[m.elems[e * nv + v] * dim + a] / nv;
    std::vector<int64_t> ids(ne);
    std::iota(ids.begin(), ids.end(), 0);
    rcb(cen, dim, ids, 0, ne, 0, nranks, elem_rank);
}

// send / receive offsets and the canonical-order reduction lists from nbr_rank / nbr_dofs
static void finish_exchange_lists(Partition& P) {
    const int rank = P.rank;
    P.nbr_offset.assign(P.nbr_rank.size() + 1, 0);
    for (size_t k = 0; k < P.nbr_rank.size(); ++k) {
        P.nbr_offset[k + 1] = P.nbr_offset[k] + (int)P.nbr_dofs[k].size();
        P.send_idx.insert(P.send_idx.end(), P.nbr_dofs[k].begin(), P.nbr_dofs[k].end());
    }
    // canonical-order reduction lists
    std::vector<std::vector<std::pair<int, int>>> contrib(P.n_local);  // (rank, recv index)
    for (size_t k = 0; k < P.nbr_rank.size(); ++k)
        for (size_t j = 0; j < P.nbr_dofs[k].size(); ++j)
            contrib[P.nbr_dofs[k][j]].emplace_back(P.nbr_rank[k], P.nbr_offset[k] + (int)j);
    P.red_ptr.push_back(0);
    for (int64_t i = 0; i < P.n_local; ++i) {
        if (contrib[i].empty()) continue;
        contrib[i].emplace_back(rank, -1);
        std::sort(contrib[i].begin(), contrib[i].end());
        P.red_dof.push_back((int)i);
        for (auto& c : contrib[i]) P.red_src.push_back(c.second);
        P.red_ptr.push_back((int)P.red_src.size());
    }
}

Partition make_partition(const Space& G, const std::vector<int>& elem_rank, int nranks, int rank) {
    if (nranks > 64) throw std::runtime_error("at most 64 subdomains are supported");
    Partition P;
    P.rank = rank; P.nranks = nranks; P.n_global = G.ndof;
    const int nloc = G.nloc;
    // ranks touching each global DOF
    std::vector<uint64_t> mask(G.ndof, 0);
    for (int64_t e = 0; e < G.nelem; ++e) {
        const uint64_t bit = 1ull << elem_rank[e];
        for (int l = 0; l < nloc; ++l) mask[G.gather[e * nloc + l]] |= bit;
    }
    const uint64_t mybit = 1ull << rank;
    // local numbering: owned first (ascending global id), then non-owned
    std::vector<int> g2l(G.ndof, -1);
    std::vector<int64_t> owned, ghost;
    for (int64_t g = 0; g < G.ndof; ++g) {
        if (!(mask[g] & mybit)) continue;
        const int owner = __builtin_ctzll(mask[g]);
        (owner == rank ? owned : ghost).push_back(g);
    }
    P.n_owned = (int64_t)owned.size();
    P.n_local = P.n_owned + (int64_t)ghost.size();
    P.global_id = owned;
    P.global_id.insert(P.global_id.end(), ghost.begin(), ghost.end());
    for (int64_t i = 0; i < P.n_local; ++i) g2l[P.global_id[i]] = (int)i;
    // local space
    Space& L = P.local;
    L.dim = G.dim; L.p = G.p; L.D = G.D; L.Q = G.Q; L.nloc = G.nloc; L.nfl = G.nfl;
    L.nodes = G.nodes; L.attributes = G.attributes;
    L.ndof = P.n_local;
    const int nv = 1 << G.dim;
    std::vector<int64_t> e_g2l(G.nelem, -1);
    for (int64_t e = 0; e < G.nelem; ++e) {
        if (elem_rank[e] != rank) continue;
        e_g2l[e] = (int64_t)P.local_elems.size();
        P.local_elems.push_back(e);
        for (int l = 0; l < nloc; ++l) L.gather.push_back(g2l[G.gather[e * nloc + l]]);
        L.elem_corner.insert(L.elem_corner.end(), G.elem_corner.begin() + e * nv * G.dim, G.elem_corner.begin() + (e + 1) * nv * G.dim);
    }
    L.nelem = (int64_t)P.local_elems.size();
    const int nfv = 1 << (G.dim - 1);
    for (int64_t b = 0; b < G.nbdr; ++b) {
        const int64_t e = G.bdr_elem[b];
        if (elem_rank[e] != rank) continue;
        L.bdr_elem.push_back(e_g2l[e]);
        L.bdr_face.push_back(G.bdr_face[b]);
        L.bdr_attr.push_back(G.bdr_attr[b]);
        for (int i = 0; i < G.nfl; ++i) L.bdr_gather.push_back(g2l[G.bdr_gather[b * G.nfl + i]]);
        L.bdr_corner.insert(L.bdr_corner.end(), G.bdr_corner.begin() + b * nfv * G.dim, G.bdr_corner.begin() + (b + 1) * nfv * G.dim);
    }
    L.nbdr = (int64_t)L.bdr_attr.size();
    // essential / Dirichlet index sets come from the GLOBAL boundary (a shared DOF may sit on a
    // boundary face that belongs to a neighbour's element)
    L.attr_dofs.assign(G.attributes.size(), {});
    for (size_t ai = 0; ai < G.attributes.size(); ++ai) {
        std::vector<int> one{G.attributes[ai]}, gl;
        G.essential_dofs(one, gl);
        for (int g : gl) if (g2l[g] >= 0) L.attr_dofs[ai].push_back(g2l[g]);
        std::sort(L.attr_dofs[ai].begin(), L.attr_dofs[ai].end());
    }
    L.has_attr_dofs = true;
    // neighbours and shared lists (ascending global id)
    std::vector<std::vector<int>> by_rank(nranks);
    for (int64_t i = 0; i < P.n_local; ++i) {
        uint64_t m = mask[P.global_id[i]] & ~mybit;
        while (m) { const int s = __builtin_ctzll(m); m &= m - 1; by_rank[s].push_back((int)i); }
    }
    for (int s = 0; s < nranks; ++s) {
        if (by_rank[s].empty()) continue;
        std::sort(by_rank[s].begin(), by_rank[s].end(), [&](int a, int b) { return P.global_id[a] < P.global_id[b]; });
        P.nbr_rank.push_back(s);
        P.nbr_dofs.push_back(by_rank[s]);
    }
    finish_exchange_lists(P);
    return P;
}

// A subdomain from the CALLER's decomposition (what ParMesh(comm, mesh) + ParFiniteElementSpace hold on one rank,
// /root/reference/src/jots_driver.cpp:178,196): local tables in the caller's L-vector numbering, the owned true DOFs in the
// caller's true-DOF order, and per neighbour the shared L-vector DOFs.  The subdomain is renumbered like make_partition
// numbers it -- owned DOFs first, in the caller's true-DOF order, so a true-DOF vector IS the leading part of a local
// vector; ghosts after them in ascending global id -- and shared lists are sorted by global id, so both sides of an
// interface agree whatever order the caller listed them in.
Partition make_partition_from_tables(int dim, int p, int64_t n_ldof, int64_t n_elem, const int* gather, const double* elem_corner,
                                     int64_t n_bdr, const int* bdr_gather, const double* bdr_corner, const int* bdr_attr,
                                     const int64_t* ldof_global_id, int64_t n_true, const int* ldof_of_true,
                                     int n_nbr, const int* nbr_rank, const int64_t* nbr_ptr, const int* nbr_ldofs,
                                     int n_attributes, const int* attributes, int rank, int nranks) {
    if (nranks > 64) throw std::runtime_error("at most 64 subdomains are supported");
    if (p < 1 || p > 4 || (dim != 2 && dim != 3)) throw std::runtime_error("partition tables: dim must be 2 or 3, order 1..4");
    Partition P;
    P.rank = rank; P.nranks = nranks;
    P.n_local = n_ldof; P.n_owned = n_true;
    // caller L index -> local index
    std::vector<int> perm((size_t)n_ldof, -1);
    for (int64_t t = 0; t < n_true; ++t) {
        const int l = ldof_of_true[t];
        if (l < 0 || l >= n_ldof || perm[l] >= 0) throw std::runtime_error("partition tables: ldof_of_true is not an injective map into the L-vector");
        perm[l] = (int)t;
    }
    std::vector<int> ghosts;
    for (int64_t l = 0; l < n_ldof; ++l) if (perm[l] < 0) ghosts.push_back((int)l);
    std::sort(ghosts.begin(), ghosts.end(), [&](int a, int b) { return ldof_global_id[a] < ldof_global_id[b]; });
    for (size_t g = 0; g < ghosts.size(); ++g) perm[ghosts[g]] = (int)(n_true + (int64_t)g);
    P.global_id.assign((size_t)n_ldof, 0);
    int64_t gmax = -1;
    for (int64_t l = 0; l < n_ldof; ++l) { P.global_id[perm[l]] = ldof_global_id[l]; gmax = std::max(gmax, ldof_global_id[l]); }
    P.n_global = gmax + 1;          // a lower bound: only used for reporting
    Space& L = P.local;
    L.dim = dim; L.p = p; L.D = p + 1; L.Q = domain_quad_points(dim, p);
    L.nloc = 1; for (int a = 0; a < dim; ++a) L.nloc *= L.D;
    L.nfl = L.nloc / L.D;
    L.nodes = gauss_lobatto_01(L.D);
    L.ndof = n_ldof; L.nelem = n_elem; L.nbdr = n_bdr;
    L.gather.resize((size_t)n_elem * L.nloc);
    for (size_t i = 0; i < L.gather.size(); ++i) {
        if (gather[i] < 0 || gather[i] >= n_ldof) throw std::runtime_error("partition tables: gather entry out of range");
        L.gather[i] = perm[gather[i]];
    }
    const int nv = 1 << dim, nfv = 1 << (dim - 1);
    L.elem_corner.assign(elem_corner, elem_corner + (size_t)n_elem * nv * dim);
    L.bdr_gather.resize((size_t)n_bdr * L.nfl);
    for (size_t i = 0; i < L.bdr_gather.size(); ++i) L.bdr_gather[i] = perm[bdr_gather[i]];
    L.bdr_corner.assign(bdr_corner, bdr_corner + (size_t)n_bdr * nfv * dim);
    L.bdr_attr.assign(bdr_attr, bdr_attr + n_bdr);
    L.bdr_elem.assign((size_t)n_bdr, -1);
    L.bdr_face.assign((size_t)n_bdr, -1);
    // the GLOBAL attribute list (ParMesh::bdr_attributes): a rank need not own a face of every attribute
    if (n_attributes > 0 && attributes) L.attributes.assign(attributes, attributes + n_attributes);
    else L.attributes = L.bdr_attr;
    std::sort(L.attributes.begin(), L.attributes.end());
    L.attributes.erase(std::unique(L.attributes.begin(), L.attributes.end()), L.attributes.end());
    // a shared DOF may lie on a boundary face of a neighbour's element only: the boundary DOF sets are completed by a
    // marker exchange when the boundary conditions are set (MFEM: GetEssentialTrueDofs synchronises the marker)
    L.sync_bdr_markers = true;
    P.local_elems.resize((size_t)n_elem);
    std::iota(P.local_elems.begin(), P.local_elems.end(), 0);
    std::vector<std::pair<int, std::vector<int>>> nb;
    for (int k = 0; k < n_nbr; ++k) {
        if (nbr_rank[k] < 0 || nbr_rank[k] >= nranks || nbr_rank[k] == rank) throw std::runtime_error("partition tables: bad neighbour rank");
        std::vector<int> d;
        for (int64_t j = nbr_ptr[k]; j < nbr_ptr[k + 1]; ++j) d.push_back(perm[nbr_ldofs[j]]);
        std::sort(d.begin(), d.end(), [&](int a, int b) { return P.global_id[a] < P.global_id[b]; });
        d.erase(std::unique(d.begin(), d.end()), d.end());
        nb.emplace_back(nbr_rank[k], std::move(d));
    }
    std::sort(nb.begin(), nb.end(), [](const std::pair<int, std::vector<int>>& a, const std::pair<int, std::vector<int>>& b) { return a.first < b.first; });
    for (auto& kv : nb) { P.nbr_rank.push_back(kv.first); P.nbr_dofs.push_back(std::move(kv.second)); }
    finish_exchange_lists(P);
    return P;
}

}  // namespace jots
