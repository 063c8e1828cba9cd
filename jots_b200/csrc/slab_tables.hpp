// Host-side fill of the 1-D tables of the slab kernels.
#pragma once
#include "apply_slab.cuh"
#include "diag_face_host.hpp"

namespace jots {

// B, G, weighted and half-weighted copies; rows q >= Q/2 are written as the exact mirror images of the lower rows
// (Gauss / Gauss-Lobatto symmetry, equal to rounding in the computed tables) because the kernels read them through
// SymTab from the lower half.
template <int D, int Q>
inline void fill_slab2_tables(const TablesRT& r, Slab2Tables<D, Q>& t) {
    for (int q = 0; q < Q; ++q) {
        const bool upper = 2 * q >= Q;                       // mirrored row (the middle row of an odd rule is kept)
        const int qs = upper ? Q - 1 - q : q;
        t.W[q] = r.W[qs];
        const double half = (Q % 2 == 1 && q == Q / 2) ? 0.5 : 1.0;
        for (int i = 0; i < D; ++i) {
            const int is = upper ? D - 1 - i : i;
            const double b = r.B[qs * D + is], g = upper ? -r.G[qs * D + is] : r.G[qs * D + is], w = r.W[qs];
            t.B[q * D + i] = b; t.G[q * D + i] = g;
            t.BW[q * D + i] = w * b; t.GW[q * D + i] = w * g;
            t.BWh[q * D + i] = half * w * b; t.GWh[q * D + i] = half * w * g;
        }
    }
}

}  // namespace jots
