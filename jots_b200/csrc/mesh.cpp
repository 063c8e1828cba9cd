#include "mesh.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

#include "basis.hpp"

namespace jots {

static const int QUAD_CORNER[2][2] = {{0, 1}, {3, 2}};                          // [b][a]
static const int HEX_CORNER[2][2][2] = {{{0, 1}, {3, 2}}, {{4, 5}, {7, 6}}};    // [c][b][a]

int corner_vertex(int dim, int a, int b, int c) {
    return dim == 2 ? QUAD_CORNER[b][a] : HEX_CORNER[c][b][a];
}

std::vector<int> Mesh::bdr_attributes() const {
    std::vector<int> a(bdr_attr);
    std::sort(a.begin(), a.end());
    a.erase(std::unique(a.begin(), a.end()), a.end());
    return a;
}

int64_t fix_orientation(Mesh& m) {
    // flip negatively oriented elements (MFEM Mesh::CheckElementOrientation(true))
    const int dim = m.dim, nv = m.nvpe();
    int64_t nbad = 0;
    for (int64_t e = 0; e < m.n_elem(); ++e) {
        int64_t* el = &m.elems[e * nv];
        const double* v0 = &m.verts[el[corner_vertex(dim, 0, 0, 0)] * dim];
        double J[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int a = 0; a < dim; ++a) {
            const double* va = &m.verts[el[corner_vertex(dim, a == 0, a == 1, a == 2)] * dim];
            for (int i = 0; i < dim; ++i) J[i][a] = va[i] - v0[i];
        }
        double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                     J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
        if (det < 0) {
            ++nbad;
            if (dim == 2) std::swap(el[1], el[3]);
            else for (int i = 0; i < 4; ++i) std::swap(el[i], el[i + 4]);
        }
    }
    return nbad;
}

Mesh read_mfem_mesh(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open mesh file: " + path);
    std::vector<std::string> lines;
    std::string ln;
    while (std::getline(f, ln)) {
        size_t a = ln.find_first_not_of(" \t\r\n");
        if (a == std::string::npos) continue;
        size_t b = ln.find_last_not_of(" \t\r\n");
        ln = ln.substr(a, b - a + 1);
        if (ln[0] == '#') continue;
        lines.push_back(ln);
    }
    if (lines.empty() || lines[0].rfind("MFEM mesh v1.0", 0) != 0)
        throw std::runtime_error("not an MFEM mesh v1.0 file: " + path);
    auto section = [&](const char* name) -> size_t {
        for (size_t i = 0; i < lines.size(); ++i) if (lines[i] == name) return i + 1;
        throw std::runtime_error(std::string("mesh file has no section '") + name + "'");
    };
    Mesh m;
    size_t i = section("dimension");
    m.dim = std::stoi(lines[i]);
    if (m.dim != 2 && m.dim != 3) throw std::runtime_error("only 2-D / 3-D meshes are supported");
    const int geom = m.dim == 2 ? 3 : 5, nv = m.nvpe(), nf = m.nvpf();
    i = section("elements");
    int64_t ne = std::stoll(lines[i]);
    m.elems.resize(ne * nv); m.elem_attr.resize(ne);
    for (int64_t e = 0; e < ne; ++e) {
        std::istringstream ss(lines[i + 1 + e]);
        int attr, g; ss >> attr >> g;
        if (g != geom) throw std::runtime_error("only SQUARE (2-D) / CUBE (3-D) elements are supported");
        m.elem_attr[e] = attr;
        for (int v = 0; v < nv; ++v) ss >> m.elems[e * nv + v];
    }
    i = section("boundary");
    int64_t nb = std::stoll(lines[i]);
    m.bdr.resize(nb * nf); m.bdr_attr.resize(nb);
    for (int64_t b = 0; b < nb; ++b) {
        std::istringstream ss(lines[i + 1 + b]);
        int attr, g; ss >> attr >> g;
        m.bdr_attr[b] = attr;
        for (int v = 0; v < nf; ++v) ss >> m.bdr[b * nf + v];
    }
    i = section("vertices");
    int64_t nvert = std::stoll(lines[i]);
    int vdim = std::stoi(lines[i + 1]);
    if (vdim != m.dim) throw std::runtime_error("space dimension != mesh dimension is not supported");
    m.verts.resize(nvert * m.dim);
    for (int64_t v = 0; v < nvert; ++v) {
        std::istringstream ss(lines[i + 2 + v]);
        for (int d = 0; d < m.dim; ++d) ss >> m.verts[v * m.dim + d];
    }
    fix_orientation(m);
    return m;
}

Mesh make_brick(int64_t nx, int64_t ny, int64_t nz, double lx, double ly, double lz) {
    Mesh m; m.dim = 3;
    const int64_t vx = nx + 1, vy = ny + 1, vz = nz + 1;
    m.verts.resize(vx * vy * vz * 3);
    for (int64_t k = 0; k < vz; ++k) for (int64_t j = 0; j < vy; ++j) for (int64_t i = 0; i < vx; ++i) {
        double* v = &m.verts[(i + vx * (j + vy * k)) * 3];
        v[0] = i * (lx / nx); v[1] = j * (ly / ny); v[2] = k * (lz / nz);
    }
    auto vid = [&](int64_t i, int64_t j, int64_t k) { return i + vx * (j + vy * k); };
    m.elems.resize(nx * ny * nz * 8); m.elem_attr.assign(nx * ny * nz, 1);
    int64_t e = 0;
    for (int64_t k = 0; k < nz; ++k) for (int64_t j = 0; j < ny; ++j) for (int64_t i = 0; i < nx; ++i, ++e) {
        int64_t* el = &m.elems[e * 8];
        el[0] = vid(i, j, k); el[1] = vid(i + 1, j, k); el[2] = vid(i + 1, j + 1, k); el[3] = vid(i, j + 1, k);
        el[4] = vid(i, j, k + 1); el[5] = vid(i + 1, j, k + 1); el[6] = vid(i + 1, j + 1, k + 1); el[7] = vid(i, j + 1, k + 1);
    }
    auto push = [&](int attr, int64_t a, int64_t b, int64_t c, int64_t d) {
        m.bdr.push_back(a); m.bdr.push_back(b); m.bdr.push_back(c); m.bdr.push_back(d); m.bdr_attr.push_back(attr);
    };
    // face(attr, a, b, fn): b outer, a inner
    for (int64_t k = 0; k < nz; ++k) for (int64_t j = 0; j < ny; ++j) push(1, vid(0, j, k), vid(0, j, k + 1), vid(0, j + 1, k + 1), vid(0, j + 1, k));
    for (int64_t k = 0; k < nz; ++k) for (int64_t i = 0; i < nx; ++i) push(2, vid(i, 0, k), vid(i + 1, 0, k), vid(i + 1, 0, k + 1), vid(i, 0, k + 1));
    for (int64_t j = 0; j < ny; ++j) for (int64_t i = 0; i < nx; ++i) push(3, vid(i, j, 0), vid(i, j + 1, 0), vid(i + 1, j + 1, 0), vid(i + 1, j, 0));
    for (int64_t k = 0; k < nz; ++k) for (int64_t j = 0; j < ny; ++j) push(4, vid(nx, j, k), vid(nx, j + 1, k), vid(nx, j + 1, k + 1), vid(nx, j, k + 1));
    for (int64_t k = 0; k < nz; ++k) for (int64_t i = 0; i < nx; ++i) push(5, vid(i, ny, k), vid(i, ny, k + 1), vid(i + 1, ny, k + 1), vid(i + 1, ny, k));
    for (int64_t j = 0; j < ny; ++j) for (int64_t i = 0; i < nx; ++i) push(6, vid(i, j, nz), vid(i + 1, j, nz), vid(i + 1, j + 1, nz), vid(i, j + 1, nz));
    m.is_brick = true;
    m.bn[0] = nx; m.bn[1] = ny; m.bn[2] = nz; m.bl[0] = lx; m.bl[1] = ly; m.bl[2] = lz;
    return m;
}

Mesh make_rect(int64_t nx, int64_t ny, double lx, double ly, double x0, double y0) {
    // 2-D rectangle, attributes 1=bottom 2=right 3=top 4=left
    Mesh m; m.dim = 2;
    const int64_t vx = nx + 1, vy = ny + 1;
    m.verts.resize(vx * vy * 2);
    for (int64_t j = 0; j < vy; ++j) for (int64_t i = 0; i < vx; ++i) {
        m.verts[(i + vx * j) * 2] = x0 + i * (lx / nx); m.verts[(i + vx * j) * 2 + 1] = y0 + j * (ly / ny);
    }
    auto vid = [&](int64_t i, int64_t j) { return i + vx * j; };
    for (int64_t j = 0; j < ny; ++j) for (int64_t i = 0; i < nx; ++i) {
        m.elems.push_back(vid(i, j)); m.elems.push_back(vid(i + 1, j)); m.elems.push_back(vid(i + 1, j + 1)); m.elems.push_back(vid(i, j + 1));
        m.elem_attr.push_back(1);
    }
    for (int64_t i = 0; i < nx; ++i) {
        m.bdr.push_back(vid(i, 0)); m.bdr.push_back(vid(i + 1, 0)); m.bdr_attr.push_back(1);
        m.bdr.push_back(vid(i + 1, ny)); m.bdr.push_back(vid(i, ny)); m.bdr_attr.push_back(3);
    }
    for (int64_t j = 0; j < ny; ++j) {
        m.bdr.push_back(vid(nx, j)); m.bdr.push_back(vid(nx, j + 1)); m.bdr_attr.push_back(2);
        m.bdr.push_back(vid(0, j + 1)); m.bdr.push_back(vid(0, j)); m.bdr_attr.push_back(4);
    }
    return m;
}

// ------------------------------------------------------------------------------------------------
// lattice numbering
// ------------------------------------------------------------------------------------------------
struct Key {
    int64_t k[5];
    bool operator==(const Key& o) const { return k[0] == o.k[0] && k[1] == o.k[1] && k[2] == o.k[2] && k[3] == o.k[3] && k[4] == o.k[4]; }
};
struct KeyHash {
    size_t operator()(const Key& a) const {
        uint64_t h = 0x9E3779B97F4A7C15ull;
        for (int i = 0; i < 5; ++i) {
            uint64_t z = (uint64_t)a.k[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            h ^= z ^ (z >> 31);
        }
        return (size_t)h;
    }
};

// orientation independent key of local lattice node l of element e
static Key node_key(const Mesh& m, int64_t e, int p, int l) {
    const int dim = m.dim, D = p + 1, nv = m.nvpe();
    const int64_t* el = &m.elems[e * nv];
    int idx[3] = {0, 0, 0}, bits[3] = {0, 0, 0};
    int ll = l;
    for (int a = 0; a < dim; ++a) { idx[a] = ll % D; ll /= D; }
    int freeax[3], nfree = 0;
    for (int a = 0; a < dim; ++a) {
        if (idx[a] > 0 && idx[a] < p) freeax[nfree++] = a;
        bits[a] = idx[a] == 0 ? 0 : 1;
    }
    Key k{{0, 0, 0, 0, 0}};
    if (nfree == 0) {
        k.k[0] = 0; k.k[1] = el[corner_vertex(dim, bits[0], bits[1], bits[2])];
    } else if (nfree == dim) {
        k.k[0] = 3; k.k[1] = e; k.k[4] = l;
    } else if (nfree == 1) {
        const int a = freeax[0];
        int b0[3] = {bits[0], bits[1], bits[2]}, b1[3] = {bits[0], bits[1], bits[2]};
        b0[a] = 0; b1[a] = 1;
        int64_t g0 = el[corner_vertex(dim, b0[0], b0[1], b0[2])], g1 = el[corner_vertex(dim, b1[0], b1[1], b1[2])];
        const int t = idx[a];
        const bool fwd = g0 < g1;
        k.k[0] = 1; k.k[1] = fwd ? g0 : g1; k.k[2] = fwd ? g1 : g0; k.k[4] = fwd ? t : p - t;
    } else {  // face of a hex: two free axes
        const int a = freeax[0], b = freeax[1];
        const int s = idx[a], t = idx[b];
        int64_t c[2][2];
        for (int ca = 0; ca < 2; ++ca) for (int cb = 0; cb < 2; ++cb) {
            int bb[3] = {bits[0], bits[1], bits[2]};
            bb[a] = ca; bb[b] = cb;
            c[ca][cb] = el[corner_vertex(dim, bb[0], bb[1], bb[2])];
        }
        int ca = 0, cb = 0;
        int64_t vmin = c[0][0];
        for (int x = 0; x < 2; ++x) for (int y = 0; y < 2; ++y) if (c[x][y] < vmin) { vmin = c[x][y]; ca = x; cb = y; }
        const int64_t na = c[1 - ca][cb], nb = c[ca][1 - cb];
        const int sp = ca == 0 ? s : p - s, tp = cb == 0 ? t : p - t;
        const bool a_first = na < nb;
        k.k[0] = 2; k.k[1] = vmin; k.k[2] = a_first ? na : nb; k.k[3] = a_first ? nb : na;
        k.k[4] = a_first ? sp * D + tp : tp * D + sp;
    }
    return k;
}

void lattice_numbering(const Mesh& m, int p, std::vector<int64_t>& gather, int64_t& nnodes) {
    const int dim = m.dim, D = p + 1;
    int nloc = 1; for (int a = 0; a < dim; ++a) nloc *= D;
    const int64_t ne = m.n_elem();
    gather.resize(ne * nloc);
    std::unordered_map<Key, int64_t, KeyHash> table;
    table.reserve((size_t)(ne * nloc / 2 + 16));
    int64_t next = 0;
    for (int64_t e = 0; e < ne; ++e)
        for (int l = 0; l < nloc; ++l) {
            Key k = node_key(m, e, p, l);
            if (k.k[0] == 3) { gather[e * nloc + l] = next++; continue; }  // element interior: always new
            auto it = table.find(k);
            if (it == table.end()) { table.emplace(k, next); gather[e * nloc + l] = next++; }
            else gather[e * nloc + l] = it->second;
        }
    nnodes = next;
}

void brick_lattice_numbering(int64_t nx, int64_t ny, int64_t nz, int p, std::vector<int64_t>& gather, int64_t& nnodes) {
    const int D = p + 1;
    const int64_t NX = nx * p + 1, NY = ny * p + 1;
    const int nloc = D * D * D;
    gather.resize(nx * ny * nz * nloc);
    int64_t e = 0;
    for (int64_t k = 0; k < nz; ++k) for (int64_t j = 0; j < ny; ++j) for (int64_t i = 0; i < nx; ++i, ++e) {
        const int64_t base = i * p + NX * (j * p + NY * (k * p));
        for (int c = 0; c < D; ++c) for (int b = 0; b < D; ++b) for (int a = 0; a < D; ++a)
            gather[e * nloc + a + D * (b + D * c)] = base + a + NX * (b + NY * c);
    }
    nnodes = NX * NY * (nz * p + 1);
}

void face_lattice_nodes(int dim, int p, int f, std::vector<int>& out) {
    const int D = p + 1, axis = f / 2, side = f % 2;
    int nfl = 1; for (int a = 0; a < dim - 1; ++a) nfl *= D;
    int pw[3] = {1, D, D * D};
    out.resize(nfl);
    for (int mth = 0; mth < nfl; ++mth) {
        int idx[3] = {0, 0, 0};
        idx[axis] = side ? p : 0;
        int mm = mth;
        for (int a = 0; a < dim; ++a) { if (a == axis) continue; idx[a] = mm % D; mm /= D; }
        int l = 0; for (int a = 0; a < dim; ++a) l += idx[a] * pw[a];
        out[mth] = l;
    }
}

void face_vertices_local(int dim, int f, std::vector<int>& out) {
    const int axis = f / 2, side = f % 2;
    const int n = 1 << (dim - 1);
    out.resize(n);
    for (int mth = 0; mth < n; ++mth) {
        int bits[3] = {0, 0, 0};
        bits[axis] = side;
        int nfree = 0;
        for (int a = 0; a < dim; ++a) { if (a == axis) continue; bits[a] = (mth >> nfree) & 1; ++nfree; }
        out[mth] = corner_vertex(dim, bits[0], bits[1], bits[2]);
    }
}

void locate_boundary_faces(const Mesh& m, std::vector<int64_t>& be, std::vector<int>& bf) {
    const int dim = m.dim, nf = 2 * dim, nv = m.nvpe(), nfv = m.nvpf();
    typedef std::array<int64_t, 4> FK;
    struct FKHash { size_t operator()(const FK& a) const {
        uint64_t h = 1469598103934665603ull;
        for (int i = 0; i < 4; ++i) { h ^= (uint64_t)a[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); }
        return (size_t)h; } };
    // only boundary faces are needed: first hash the boundary elements, then look the element faces up
    std::unordered_map<FK, int64_t, FKHash> wanted;
    wanted.reserve((size_t)m.n_bdr() * 2 + 16);
    for (int64_t b = 0; b < m.n_bdr(); ++b) {
        FK k{{-1, -1, -1, -1}};
        for (int v = 0; v < nfv; ++v) k[v] = m.bdr[b * nfv + v];
        std::sort(k.begin(), k.begin() + nfv);
        wanted.emplace(k, -1);
    }
    std::unordered_map<FK, std::pair<int64_t, int>, FKHash> found;
    found.reserve(wanted.size() * 2 + 16);
    std::vector<int> lv;
    // oracle order: f outer loop, e inner loop, first hit wins
    std::vector<std::vector<int>> lvs(nf);
    for (int f = 0; f < nf; ++f) face_vertices_local(dim, f, lvs[f]);
    // cheap filter before the hash: a face can only be a boundary face if every vertex of it lies on some boundary face
    std::vector<unsigned char> on_bdr((size_t)m.n_verts(), 0);
    for (size_t i = 0; i < m.bdr.size(); ++i) on_bdr[(size_t)m.bdr[i]] = 1;
    for (int f = 0; f < nf; ++f)
        for (int64_t e = 0; e < m.n_elem(); ++e) {
            FK k{{-1, -1, -1, -1}};
            bool all = true;
            for (int v = 0; v < nfv; ++v) { k[v] = m.elems[e * nv + lvs[f][v]]; all = all && on_bdr[(size_t)k[v]]; }
            if (!all) continue;
            std::sort(k.begin(), k.begin() + nfv);
            if (wanted.find(k) == wanted.end()) continue;
            found.emplace(k, std::make_pair(e, f));
        }
    be.resize(m.n_bdr()); bf.resize(m.n_bdr());
    for (int64_t b = 0; b < m.n_bdr(); ++b) {
        FK k{{-1, -1, -1, -1}};
        for (int v = 0; v < nfv; ++v) k[v] = m.bdr[b * nfv + v];
        std::sort(k.begin(), k.begin() + nfv);
        auto it = found.find(k);
        if (it == found.end()) throw std::runtime_error("boundary element does not match any element face");
        be[b] = it->second.first; bf[b] = it->second.second;
    }
}

// ------------------------------------------------------------------------------------------------
// uniform refinement
// ------------------------------------------------------------------------------------------------
Mesh uniform_refine(const Mesh& m) {
    const int dim = m.dim, nv = m.nvpe();
    std::vector<int64_t> gather; int64_t nn;
    lattice_numbering(m, 2, gather, nn);
    const int nloc = dim == 2 ? 9 : 27;
    Mesh out; out.dim = dim;
    out.verts.assign(nn * dim, 0.0);
    const double xi[3] = {0.0, 0.5, 1.0};
    const int64_t ne = m.n_elem();
    for (int64_t e = 0; e < ne; ++e)
        for (int l = 0; l < nloc; ++l) {
            int idx[3] = {l % 3, (l / 3) % 3, dim == 3 ? l / 9 : 0};
            double pt[3] = {0, 0, 0};
            for (int c = 0; c < (1 << dim); ++c) {
                int cb[3] = {c & 1, (c >> 1) & 1, (c >> 2) & 1};
                double w = 1.0;
                for (int a = 0; a < dim; ++a) w *= cb[a] ? xi[idx[a]] : 1.0 - xi[idx[a]];
                if (w == 0.0) continue;
                const double* X = &m.verts[m.elems[e * nv + corner_vertex(dim, cb[0], cb[1], cb[2])] * dim];
                for (int i = 0; i < dim; ++i) pt[i] += w * X[i];
            }
            for (int i = 0; i < dim; ++i) out.verts[gather[e * nloc + l] * dim + i] = pt[i];
        }
    // children: sub = (s0, s1[, s2]) with the LAST axis fastest
    const int nch = 1 << dim;
    out.elems.resize(ne * nch * nv); out.elem_attr.resize(ne * nch);
    for (int64_t e = 0; e < ne; ++e)
        for (int ch = 0; ch < nch; ++ch) {
            int sub[3] = {0, 0, 0};
            if (dim == 2) { sub[0] = (ch >> 1) & 1; sub[1] = ch & 1; }
            else { sub[0] = (ch >> 2) & 1; sub[1] = (ch >> 1) & 1; sub[2] = ch & 1; }
            int64_t* child = &out.elems[(e * nch + ch) * nv];
            for (int c = 0; c < (1 << dim); ++c) {
                int cb[3] = {c & 1, (c >> 1) & 1, (c >> 2) & 1};
                int l = 0, pw = 1;
                for (int a = 0; a < dim; ++a) { l += (sub[a] + cb[a]) * pw; pw *= 3; }
                child[corner_vertex(dim, cb[0], cb[1], cb[2])] = gather[e * nloc + l];
            }
            out.elem_attr[e * nch + ch] = m.elem_attr[e];
        }
    // boundary
    std::vector<int64_t> be; std::vector<int> bf;
    locate_boundary_faces(m, be, bf);
    const int64_t nb = m.n_bdr();
    std::vector<int> nodes;
    if (dim == 2) {
        out.bdr.resize(nb * 2 * 2); out.bdr_attr.resize(nb * 2);
        for (int64_t b = 0; b < nb; ++b) {
            face_lattice_nodes(2, 2, bf[b], nodes);
            const int64_t* g = &gather[be[b] * nloc];
            out.bdr[(b * 2 + 0) * 2 + 0] = g[nodes[0]]; out.bdr[(b * 2 + 0) * 2 + 1] = g[nodes[1]];
            out.bdr[(b * 2 + 1) * 2 + 0] = g[nodes[1]]; out.bdr[(b * 2 + 1) * 2 + 1] = g[nodes[2]];
            out.bdr_attr[b * 2] = out.bdr_attr[b * 2 + 1] = m.bdr_attr[b];
        }
    } else {
        out.bdr.resize(nb * 4 * 4); out.bdr_attr.resize(nb * 4);
        for (int64_t b = 0; b < nb; ++b) {
            face_lattice_nodes(3, 2, bf[b], nodes);  // [t][s], s fastest
            const int64_t* g = &gather[be[b] * nloc];
            auto G = [&](int t, int s) { return g[nodes[t * 3 + s]]; };
            int c = 0;
            for (int tt = 0; tt < 2; ++tt) for (int ss = 0; ss < 2; ++ss, ++c) {
                int64_t* q = &out.bdr[(b * 4 + c) * 4];
                q[0] = G(tt, ss); q[1] = G(tt, ss + 1); q[2] = G(tt + 1, ss + 1); q[3] = G(tt + 1, ss);
                out.bdr_attr[b * 4 + c] = m.bdr_attr[b];
            }
        }
    }
    fix_orientation(out);
    return out;
}

// ------------------------------------------------------------------------------------------------
// space
// ------------------------------------------------------------------------------------------------
Space make_space(const Mesh& m, int p) {
    if (p < 1 || p > 4) throw std::runtime_error("FE_Order must be 1..4");
    Space s;
    s.dim = m.dim; s.p = p; s.D = p + 1; s.Q = domain_quad_points(m.dim, p);
    s.nloc = 1; for (int a = 0; a < m.dim; ++a) s.nloc *= s.D;
    s.nfl = s.nloc / s.D;
    s.nelem = m.n_elem(); s.nbdr = m.n_bdr();
    s.nodes = gauss_lobatto_01(s.D);
    s.is_brick = m.is_brick;
    for (int a = 0; a < 3; ++a) { s.bn[a] = m.bn[a]; s.bl[a] = m.bl[a]; }
    std::vector<int64_t> g64; int64_t nn;
    if (m.is_brick) brick_lattice_numbering(m.bn[0], m.bn[1], m.bn[2], p, g64, nn);
    else lattice_numbering(m, p, g64, nn);
    if (nn >= (int64_t)1 << 31) throw std::runtime_error("more than 2^31-1 DOFs on one device are not supported");
    s.ndof = nn;
    s.gather.resize(g64.size());
    for (size_t i = 0; i < g64.size(); ++i) s.gather[i] = (int)g64[i];
    const int nv = m.nvpe(), dim = m.dim;
    s.elem_corner.resize(s.nelem * nv * dim);
    for (int64_t e = 0; e < s.nelem; ++e)
        for (int c = 0; c < nv; ++c) {
            int64_t v = m.elems[e * nv + corner_vertex(dim, c & 1, (c >> 1) & 1, (c >> 2) & 1)];
            for (int i = 0; i < dim; ++i) s.elem_corner[(e * nv + c) * dim + i] = m.verts[v * dim + i];
        }
    locate_boundary_faces(m, s.bdr_elem, s.bdr_face);
    s.bdr_attr = m.bdr_attr;
    s.attributes = m.bdr_attributes();
    s.bdr_gather.resize(s.nbdr * s.nfl);
    const int nfv = m.nvpf();
    s.bdr_corner.resize(s.nbdr * nfv * dim);
    std::vector<std::vector<int>> fl(2 * dim), fv(2 * dim);
    for (int f = 0; f < 2 * dim; ++f) { face_lattice_nodes(dim, p, f, fl[f]); face_vertices_local(dim, f, fv[f]); }
    for (int64_t b = 0; b < s.nbdr; ++b) {
        const int f = s.bdr_face[b];
        const int64_t e = s.bdr_elem[b];
        for (int i = 0; i < s.nfl; ++i) s.bdr_gather[b * s.nfl + i] = s.gather[e * s.nloc + fl[f][i]];
        for (int c = 0; c < nfv; ++c) {
            int64_t v = m.elems[e * nv + fv[f][c]];
            for (int i = 0; i < dim; ++i) s.bdr_corner[(b * nfv + c) * dim + i] = m.verts[v * dim + i];
        }
    }
    return s;
}

void Space::node_coordinates(std::vector<double>& out) const {
    out.assign(ndof * dim, 0.0);
    const int nv = 1 << dim;
    for (int64_t e = 0; e < nelem; ++e)
        for (int l = 0; l < nloc; ++l) {
            int idx[3] = {l % D, (l / D) % D, dim == 3 ? l / (D * D) : 0};
            double pt[3] = {0, 0, 0};
            for (int c = 0; c < nv; ++c) {
                double w = 1.0;
                for (int a = 0; a < dim; ++a) w *= ((c >> a) & 1) ? nodes[idx[a]] : 1.0 - nodes[idx[a]];
                for (int i = 0; i < dim; ++i) pt[i] += w * elem_corner[(e * nv + c) * dim + i];
            }
            for (int i = 0; i < dim; ++i) out[(int64_t)gather[e * nloc + l] * dim + i] = pt[i];
        }
}

void Space::essential_dofs(const std::vector<int>& attrs, std::vector<int>& out) const {
    out.clear();
    if (has_attr_dofs) {
        for (size_t ai = 0; ai < attributes.size(); ++ai)
            if (std::find(attrs.begin(), attrs.end(), attributes[ai]) != attrs.end())
                out.insert(out.end(), attr_dofs[ai].begin(), attr_dofs[ai].end());
        std::sort(out.begin(), out.end());
        out.erase(std::unique(out.begin(), out.end()), out.end());
        return;
    }
    for (int64_t b = 0; b < nbdr; ++b) {
        if (std::find(attrs.begin(), attrs.end(), bdr_attr[b]) == attrs.end()) continue;
        for (int i = 0; i < nfl; ++i) out.push_back(bdr_gather[b * nfl + i]);
    }
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
}

void h1_dof_map(int dim, int p, std::vector<int>& map) {
    if ((dim != 2 && dim != 3) || p < 1) throw std::runtime_error("h1_dof_map: dim 2 or 3, order >= 1");
    const int p1 = p + 1;
    int o = 0;
    if (dim == 2) {
        map.assign((size_t)p1 * p1, -1);
        auto at = [&](int i, int j) -> int& { return map[(size_t)i + (size_t)j * p1]; };
        at(0, 0) = o++; at(p, 0) = o++; at(p, p) = o++; at(0, p) = o++;
        for (int i = 1; i < p; ++i) at(i, 0) = o++;            // edge (0,1)
        for (int i = 1; i < p; ++i) at(p, i) = o++;            // edge (1,2)
        for (int i = 1; i < p; ++i) at(p - i, p) = o++;        // edge (2,3)
        for (int i = 1; i < p; ++i) at(0, p - i) = o++;        // edge (3,0)
        for (int j = 1; j < p; ++j) for (int i = 1; i < p; ++i) at(i, j) = o++;
        return;
    }
    map.assign((size_t)p1 * p1 * p1, -1);
    auto at = [&](int i, int j, int k) -> int& { return map[(size_t)i + ((size_t)j + (size_t)k * p1) * p1]; };
    at(0, 0, 0) = o++; at(p, 0, 0) = o++; at(p, p, 0) = o++; at(0, p, 0) = o++;
    at(0, 0, p) = o++; at(p, 0, p) = o++; at(p, p, p) = o++; at(0, p, p) = o++;
    // edges, in the order of Hexahedron::edges: (0,1) (1,2) (3,2) (0,3) (4,5) (5,6) (7,6) (4,7) (0,4) (1,5) (2,6) (3,7)
    for (int i = 1; i < p; ++i) at(i, 0, 0) = o++;
    for (int i = 1; i < p; ++i) at(p, i, 0) = o++;
    for (int i = 1; i < p; ++i) at(i, p, 0) = o++;
    for (int i = 1; i < p; ++i) at(0, i, 0) = o++;
    for (int i = 1; i < p; ++i) at(i, 0, p) = o++;
    for (int i = 1; i < p; ++i) at(p, i, p) = o++;
    for (int i = 1; i < p; ++i) at(i, p, p) = o++;
    for (int i = 1; i < p; ++i) at(0, i, p) = o++;
    for (int i = 1; i < p; ++i) at(0, 0, i) = o++;
    for (int i = 1; i < p; ++i) at(p, 0, i) = o++;
    for (int i = 1; i < p; ++i) at(p, p, i) = o++;
    for (int i = 1; i < p; ++i) at(0, p, i) = o++;
    // faces, in the order of Mesh::GenerateFaces: (3,2,1,0) (0,1,5,4) (1,2,6,5) (2,3,7,6) (3,0,4,7) (4,5,6,7)
    for (int j = 1; j < p; ++j) for (int i = 1; i < p; ++i) at(i, p - j, 0) = o++;
    for (int j = 1; j < p; ++j) for (int i = 1; i < p; ++i) at(i, 0, j) = o++;
    for (int j = 1; j < p; ++j) for (int i = 1; i < p; ++i) at(p, i, j) = o++;
    for (int j = 1; j < p; ++j) for (int i = 1; i < p; ++i) at(p - i, p, j) = o++;
    for (int j = 1; j < p; ++j) for (int i = 1; i < p; ++i) at(0, p - i, j) = o++;
    for (int j = 1; j < p; ++j) for (int i = 1; i < p; ++i) at(i, j, p) = o++;
    for (int k = 1; k < p; ++k) for (int j = 1; j < p; ++j) for (int i = 1; i < p; ++i) at(i, j, k) = o++;
}

}  // namespace jots

namespace jots {

void fill_boundary_adjacency(Space& s) {
    if ((int64_t)s.bdr_elem.size() != s.nbdr) { s.bdr_elem.assign((size_t)s.nbdr, -1); s.bdr_face.assign((size_t)s.nbdr, -1); }
    bool missing = false;
    for (int64_t b = 0; b < s.nbdr; ++b) missing = missing || s.bdr_elem[b] < 0;
    if (!missing) return;
    const int dim = s.dim, p = s.p, D = s.D, nfv = 1 << (dim - 1);
    // corner DOFs (sorted) of a face lattice given as nfl node ids, lower free axis fastest
    auto corners = [&](const int* nodes, std::array<int, 4>& key) {
        key = {-1, -1, -1, -1};
        if (dim == 2) { key[0] = nodes[0]; key[1] = nodes[D - 1]; }
        else { key[0] = nodes[0]; key[1] = nodes[D - 1]; key[2] = nodes[(D - 1) * D]; key[3] = nodes[D * D - 1]; }
        for (int i = 1; i < nfv; ++i)        // insertion sort of two or four ids
            for (int j = i; j > 0 && key[j - 1] > key[j]; --j) std::swap(key[j - 1], key[j]);
    };
    std::map<std::array<int, 4>, std::pair<int64_t, int>> faces;
    std::vector<int> fl, nodes((size_t)s.nfl);
    std::vector<std::vector<int>> face_nodes(2 * dim);
    for (int f = 0; f < 2 * dim; ++f) face_lattice_nodes(dim, p, f, face_nodes[f]);
    // only the faces that carry a DOF of a boundary element can match: mark those DOFs first
    std::vector<unsigned char> on_bdr((size_t)s.ndof, 0);
    for (int g : s.bdr_gather) on_bdr[g] = 1;
    for (int64_t e = 0; e < s.nelem; ++e) {
        const int* g = &s.gather[(size_t)e * s.nloc];
        for (int f = 0; f < 2 * dim; ++f) {
            const std::vector<int>& fn = face_nodes[f];
            if (!on_bdr[g[fn[0]]] || !on_bdr[g[fn[s.nfl - 1]]]) continue;
            for (int m = 0; m < s.nfl; ++m) nodes[m] = g[fn[m]];
            std::array<int, 4> key;
            corners(nodes.data(), key);
            faces[key] = {e, f};
        }
    }
    for (int64_t b = 0; b < s.nbdr; ++b) {
        if (s.bdr_elem[b] >= 0) continue;
        std::array<int, 4> key;
        corners(&s.bdr_gather[(size_t)b * s.nfl], key);
        auto it = faces.find(key);
        if (it == faces.end()) throw std::runtime_error("boundary element " + std::to_string(b) + " matches no element face (boundary DOF table inconsistent with the element table)");
        s.bdr_elem[b] = it->second.first;
        s.bdr_face[b] = it->second.second;
    }
}

void boundary_node_locals(const Space& s, std::vector<int>& loc) {
    loc.assign((size_t)s.nbdr * s.nfl, -1);
    std::vector<int> fn;
    for (int64_t b = 0; b < s.nbdr && b < (int64_t)s.bdr_elem.size(); ++b) {
        const int64_t e = s.bdr_elem[b];
        if (e < 0) continue;
        face_lattice_nodes(s.dim, s.p, s.bdr_face[b], fn);
        const int* g = &s.gather[(size_t)e * s.nloc];
        for (int m = 0; m < s.nfl; ++m) {
            const int dof = s.bdr_gather[(size_t)b * s.nfl + m];
            int found = -1;
            if (g[fn[m]] == dof) found = fn[m];                       // the usual case: same order as the element's face lattice
            else for (int q = 0; q < s.nfl; ++q) if (g[fn[q]] == dof) { found = fn[q]; break; }
            if (found < 0) throw std::runtime_error("boundary element " + std::to_string(b) + ": DOF not on the matched element face");
            loc[(size_t)b * s.nfl + m] = found;
        }
    }
}

}  // namespace jots
