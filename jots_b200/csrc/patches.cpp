// Host-side builder of the element patches of the unique-pencil slab kernel (apply_patch.cuh).
//
// A patch is a PX x PY x 1 block of lattice-adjacent hexes whose local axes are aligned; its (D-1)PX+1 by (D-1)PY+1 by D
// nodes are listed once ([k][J][I], I fastest), so a nodal pencil shared by two or four elements of the patch is gathered,
// contracted and scattered once.  Nothing is assumed about the numbering of the L-vector (lexicographic bricks, first-
// appearance numbering of file meshes and owned-first numbering of partitioned subdomains all work): the lattice is
// recovered from the gather table alone.  The reference has no counterpart (MFEM assembles element by element:
// /root/reference/src/nl_conduction_operator.cpp:41-67); this replaces the per-element gather of ParNonlinearForm::Mult.
#include <algorithm>
#include <cstring>
#include <stdexcept>

#include "mesh.hpp"

namespace jots {

namespace {

// plus-direction neighbour of every element along local axis a (-1: none), through the face-interior node the two
// elements share; the whole face must match node by node, which also checks that the local axes are aligned
void axis_neighbours(const Space& s, int axis, std::vector<int>& owner, std::vector<int>& nbr) {
    const int D = s.D, ND = D * D * D;
    const int64_t ne = s.nelem;
    auto lid = [&](int a, int u, int v, int w) {   // local index of (axis coordinate w, in-face coordinates u, v)
        int c[3];
        c[a] = w; c[(a + 1) % 3] = u; c[(a + 2) % 3] = v;
        return c[0] + D * (c[1] + D * c[2]);
    };
    std::fill(owner.begin(), owner.end(), -1);
    const int* G = s.gather.data();
    const int fm = lid(axis, 1, 1, 0), fp = lid(axis, 1, 1, D - 1);   // a face-interior node (D >= 3)
    for (int64_t e = 0; e < ne; ++e) owner[G[e * ND + fm]] = (int)e;
    nbr.assign(ne, -1);
    for (int64_t e = 0; e < ne; ++e) {
        const int o = owner[G[e * ND + fp]];
        if (o < 0 || o == e) continue;
        bool same = true;
        for (int v = 0; v < D && same; ++v)
            for (int u = 0; u < D; ++u)
                if (G[e * ND + lid(axis, u, v, D - 1)] != G[(int64_t)o * ND + lid(axis, u, v, 0)]) { same = false; break; }
        if (same) nbr[e] = o;
    }
}

}  // namespace

int64_t make_patches(const Space& s, const std::vector<int>& ess, int PX, int PY, std::vector<int>& tab, double* fill) {
    tab.clear();
    if (fill) *fill = 0.0;
    const int D = s.D, DD = D * D, ND = DD * D;
    if (s.dim != 3 || s.nelem == 0 || D < 3 || s.nloc != ND || s.ndof > 0x7fffffffLL || s.nelem > 0x7fffffffLL) return 0;
    const int64_t ne = s.nelem;
    const int NXN = (D - 1) * PX + 1, NYN = (D - 1) * PY + 1, NPEN = NXN * NYN, NN = NPEN * D, E = PX * PY;
    const int REC = (NN + E + 3) / 4 * 4;
    const int* G = s.gather.data();

    // ---- lattice coordinates of every element -------------------------------------------------------------------------
    std::vector<int> cx(ne), cy(ne), cz(ne), comp(ne, -1);
    int ncomp = 0;
    if (s.is_brick && s.bn[0] * s.bn[1] * s.bn[2] == ne) {
        // make_brick: element e = ix + nx (iy + ny iz), local axes = global axes
        int64_t e = 0;
        for (int64_t k = 0; k < s.bn[2]; ++k) for (int64_t j = 0; j < s.bn[1]; ++j) for (int64_t i = 0; i < s.bn[0]; ++i, ++e) {
            cx[e] = (int)i; cy[e] = (int)j; cz[e] = (int)k; comp[e] = 0;
        }
        ncomp = 1;
    } else {
        std::vector<int> owner((size_t)s.ndof), nb[3], pv[3];
        for (int a = 0; a < 3; ++a) {
            axis_neighbours(s, a, owner, nb[a]);
            pv[a].assign(ne, -1);
            for (int64_t e = 0; e < ne; ++e) if (nb[a][e] >= 0) pv[a][nb[a][e]] = (int)e;
        }
        owner.clear(); owner.shrink_to_fit();
        std::vector<int> queue;
        queue.reserve(ne);
        for (int64_t s0 = 0; s0 < ne; ++s0) {
            if (comp[s0] >= 0) continue;
            comp[s0] = ncomp; cx[s0] = cy[s0] = cz[s0] = 0;
            queue.clear(); queue.push_back((int)s0);
            for (size_t h = 0; h < queue.size(); ++h) {
                const int e = queue[h];
                for (int a = 0; a < 3; ++a)
                    for (int sgn = -1; sgn <= 1; sgn += 2) {
                        const int o = sgn > 0 ? nb[a][e] : pv[a][e];
                        if (o < 0) continue;
                        const int ox = cx[e] + (a == 0 ? sgn : 0), oy = cy[e] + (a == 1 ? sgn : 0), oz = cz[e] + (a == 2 ? sgn : 0);
                        if (comp[o] < 0) { comp[o] = ncomp; cx[o] = ox; cy[o] = oy; cz[o] = oz; queue.push_back(o); }
                        else if (cx[o] != ox || cy[o] != oy || cz[o] != oz) return 0;   // not a lattice (twisted / periodic)
                    }
            }
            ++ncomp;
        }
        // shift every component to non-negative coordinates
        std::vector<int> mn((size_t)ncomp * 3, 0x7fffffff);
        for (int64_t e = 0; e < ne; ++e) {
            int* m = &mn[(size_t)comp[e] * 3];
            m[0] = std::min(m[0], cx[e]); m[1] = std::min(m[1], cy[e]); m[2] = std::min(m[2], cz[e]);
        }
        for (int64_t e = 0; e < ne; ++e) {
            const int* m = &mn[(size_t)comp[e] * 3];
            cx[e] -= m[0]; cy[e] -= m[1]; cz[e] -= m[2];
        }
    }

    // ---- group by (component, layer, patch row, patch column) -------------------------------------------------------------
    struct Key { int comp, z, py, px, e; };
    std::vector<Key> keys(ne);
    for (int64_t e = 0; e < ne; ++e) keys[e] = Key{comp[e], cz[e], cy[e] / PY, cx[e] / PX, (int)e};
    std::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) {
        if (a.comp != b.comp) return a.comp < b.comp;
        if (a.z != b.z) return a.z < b.z;
        if (a.py != b.py) return a.py < b.py;
        if (a.px != b.px) return a.px < b.px;
        return a.e < b.e;
    });
    std::vector<unsigned char> em((size_t)s.ndof, 0);
    for (int d : ess) em[d] = 1;
    const int INVALID = (int)0x80000000;
    int64_t npatch = 0;
    tab.reserve((size_t)(ne / E + 16) * REC);
    for (int64_t i = 0; i < ne;) {
        int64_t j = i;
        while (j < ne && keys[j].comp == keys[i].comp && keys[j].z == keys[i].z && keys[j].py == keys[i].py && keys[j].px == keys[i].px) ++j;
        const size_t o = tab.size();
        tab.resize(o + REC, 0);
        int* rec = &tab[o];
        for (int m = 0; m < NN; ++m) rec[m] = INVALID;
        for (int m = 0; m < E; ++m) rec[NN + m] = -1;
        for (int64_t q = i; q < j; ++q) {
            const int e = keys[q].e, lx = cx[e] % PX, ly = cy[e] % PY, slot = ly * PX + lx;
            if (rec[NN + slot] >= 0) { tab.clear(); return 0; }   // two elements on one lattice site
            rec[NN + slot] = e;
            for (int k = 0; k < D; ++k) for (int jj = 0; jj < D; ++jj) for (int ii = 0; ii < D; ++ii) {
                const int g = G[(int64_t)e * ND + ii + D * (jj + D * k)];
                int& dst = rec[k * NPEN + ((D - 1) * ly + jj) * NXN + (D - 1) * lx + ii];
                const int val = em[g] ? ~g : g;
                if (dst != INVALID && dst != val) { tab.clear(); return 0; }   // neighbours disagree: not a conforming lattice
                dst = val;
            }
        }
        ++npatch;
        i = j;
    }
    if (fill) *fill = (double)ne / ((double)npatch * E);
    return npatch;
}

}  // namespace jots
