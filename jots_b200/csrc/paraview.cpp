// ParaView output of the reference's OutputManager (/root/reference/src/output_manager.cpp:23-28,81-87): what
// mfem::ParaViewDataCollection::Save writes with SetLevelsOfDetail(order), SetDataFormat(VTKFormat::BINARY) and
// SetHighOrderOutput(true) --
//
//   ParaView/ParaView.pvd                      collection: one <DataSet timestep=.. file="Cycle000010/data.pvtu"/> per save
//   ParaView/Cycle<cycle:06d>/data.pvtu        parallel header naming the arrays and the pieces
//   ParaView/Cycle<cycle:06d>/proc<rank:06d>.vtu  the piece: every element is one VTK_LAGRANGE_QUADRILATERAL (70) /
//                                              VTK_LAGRANGE_HEXAHEDRON (72) cell of the FE order with ITS OWN (p+1)^dim points
//                                              (uniform lattice of the reference element, lexicographic; the connectivity
//                                              carries VTK's vertex / edge / face / interior order), point data = the field
//                                              evaluated at those points, cell data "attribute"; inline base64 blocks, each
//                                              preceded by its own base64-encoded UInt32 byte count (no compression)
//
// restated from the VTK XML file format and MFEM 4.7's mesh/vtk.cpp + general/binaryio (neither is part of the reference
// tree): byte-level identity with MFEM's files is NOT pinned; the files are checked by an independent reader in
// tests/test_host_cpu.py (cell nodes and values against the polynomial they sample).  Single-domain form as the restart
// files: on several GPUs rank 0 writes the whole field, "Rank" holds the owner of each node.
#include "paraview.hpp"

#include <sys/stat.h>

#include <cerrno>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "basis.hpp"

namespace jots {

namespace {
// vtkLagrangeHexahedron::PointIndexFromIJK (VTK 9: the vertical edges are ordered 0-4, 1-5, 3-7, 2-6 ... as the file
// version 2.2 that MFEM writes for high-order output declares)
int vtk_hex_index(int i, int j, int k, int n) {
    const bool ib = (i == 0 || i == n), jb = (j == 0 || j == n), kb = (k == 0 || k == n);
    const int nb = (int)ib + (int)jb + (int)kb, m = n - 1;
    if (nb == 3) return (i ? (j ? 2 : 1) : (j ? 3 : 0)) + (k ? 4 : 0);
    int off = 8;
    if (nb == 2) {
        if (!ib) return (i - 1) + (j ? 2 * m : 0) + (k ? 4 * m : 0) + off;
        if (!jb) return (j - 1) + (i ? m : 3 * m) + (k ? 4 * m : 0) + off;
        off += 8 * m;
        return (k - 1) + m * (i ? (j ? 3 : 1) : (j ? 2 : 0)) + off;
    }
    off += 12 * m;
    if (nb == 1) {
        if (ib) return (j - 1) + m * (k - 1) + (i ? m * m : 0) + off;
        off += 2 * m * m;
        if (jb) return (i - 1) + m * (k - 1) + (j ? m * m : 0) + off;
        off += 2 * m * m;
        return (i - 1) + m * (j - 1) + (k ? m * m : 0) + off;
    }
    off += 6 * m * m;
    return off + (i - 1) + m * ((j - 1) + m * (k - 1));
}
// vtkLagrangeQuadrilateral::PointIndexFromIJK
int vtk_quad_index(int i, int j, int n) {
    const bool ib = (i == 0 || i == n), jb = (j == 0 || j == n);
    const int nb = (int)ib + (int)jb, m = n - 1;
    if (nb == 2) return i ? (j ? 2 : 1) : (j ? 3 : 0);
    int off = 4;
    if (nb == 1) {
        if (!ib) return (i - 1) + (j ? 2 * m : 0) + off;
        return (j - 1) + (i ? m : 3 * m) + off;
    }
    off += 4 * m;
    return off + (i - 1) + m * (j - 1);
}

void b64(std::ostream& os, const void* data, size_t n) {
    static const char T[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    const unsigned char* p = static_cast<const unsigned char*>(data);
    std::string out;
    out.reserve((n + 2) / 3 * 4);
    size_t i = 0;
    for (; i + 2 < n; i += 3) {
        const unsigned v = (p[i] << 16) | (p[i + 1] << 8) | p[i + 2];
        out.push_back(T[v >> 18]); out.push_back(T[(v >> 12) & 63]); out.push_back(T[(v >> 6) & 63]); out.push_back(T[v & 63]);
    }
    if (i + 1 == n) {
        const unsigned v = p[i] << 16;
        out.push_back(T[v >> 18]); out.push_back(T[(v >> 12) & 63]); out.push_back('='); out.push_back('=');
    } else if (i + 2 == n) {
        const unsigned v = (p[i] << 16) | (p[i + 1] << 8);
        out.push_back(T[v >> 18]); out.push_back(T[(v >> 12) & 63]); out.push_back(T[(v >> 6) & 63]); out.push_back('=');
    }
    os << out;
}
// one inline binary block: base64(UInt32 byte count) followed by base64(data)
template <class T>
void block(std::ostream& os, const std::vector<T>& v) {
    const uint64_t bytes = (uint64_t)v.size() * sizeof(T);
    if (bytes > 0xffffffffull) throw std::runtime_error("ParaView output: a data array exceeds the 4 GiB of a UInt32 block header");
    const uint32_t nb = (uint32_t)bytes;
    b64(os, &nb, sizeof(nb));
    b64(os, v.data(), (size_t)bytes);
    os << "\n";
}

void make_dir(const std::string& d) {
    if (mkdir(d.c_str(), 0777) != 0 && errno != EEXIST) throw std::runtime_error("cannot create directory " + d);
}
}  // namespace

void vtk_lagrange_order(int dim, int p, std::vector<int>& lex_of_vtk) {
    const int D = p + 1;
    int nloc = 1;
    for (int a = 0; a < dim; ++a) nloc *= D;
    lex_of_vtk.assign(nloc, -1);
    for (int k = 0; k < (dim == 3 ? D : 1); ++k)
        for (int j = 0; j < D; ++j)
            for (int i = 0; i < D; ++i) {
                const int v = dim == 3 ? vtk_hex_index(i, j, k, p) : vtk_quad_index(i, j, p);
                if (v < 0 || v >= nloc || lex_of_vtk[v] != -1) throw std::runtime_error("vtk_lagrange_order: not a permutation");
                lex_of_vtk[v] = i + D * (j + D * k);
            }
}

void ParaViewCollection::save(int cycle, double time, const Space& s, const std::vector<int>& elem_attr, const std::vector<VizField>& fields) {
    const int dim = s.dim, p = s.p, D = s.D, nloc = s.nloc, nc = 1 << dim;
    const int64_t ne = s.nelem, np = ne * nloc;
    if (np > 0x7fffffffll) throw std::runtime_error("ParaView output: more than 2^31 points in one piece (Int32 connectivity)");
    // ---- collection file -------------------------------------------------------------------------------------------------
    make_dir(dir);
    const std::string pvd = dir + "/" + name + ".pvd";
    if (!opened) {
        entries.clear();
        if (restart_mode || append) {
            // UseRestartMode: keep what an earlier run saved before the current time
            std::ifstream in(pvd);
            std::string line;
            while (std::getline(in, line)) {
                const size_t a = line.find("timestep=\""), b = line.find("file=\"");
                if (a == std::string::npos || b == std::string::npos) continue;
                const double t = std::atof(line.c_str() + a + 10);
                const size_t e = line.find('"', b + 6);
                if ((append || t < time) && e != std::string::npos) entries.push_back({t, line.substr(b + 6, e - (b + 6))});
            }
        }
        opened = true;
    }
    char cyc[32];
    std::snprintf(cyc, sizeof(cyc), "Cycle%06d", cycle);
    const std::string cdir = dir + "/" + cyc;
    make_dir(cdir);
    entries.push_back({time, std::string(cyc) + "/data.pvtu"});
    {
        std::ofstream os(pvd);
        os.precision(16);
        os << "<?xml version=\"1.0\"?>\n<VTKFile type=\"Collection\" version=\"0.1\" byte_order=\"LittleEndian\">\n<Collection>\n";
        for (const auto& e : entries)
            os << "<DataSet timestep=\"" << e.first << "\" group=\"\" part=\"0\" file=\"" << e.second << "\" name=\"mesh\"/>\n";
        os << "</Collection>\n</VTKFile>\n";
        if (!os) throw std::runtime_error("cannot write " + pvd);
    }
    // ---- parallel header ---------------------------------------------------------------------------------------------------
    {
        std::ofstream os(cdir + "/data.pvtu");
        os << "<?xml version=\"1.0\"?>\n<VTKFile type=\"PUnstructuredGrid\" version =\"2.2\" byte_order=\"LittleEndian\">\n"
           << "<PUnstructuredGrid GhostLevel=\"0\">\n<PPoints>\n\t<PDataArray type=\"Float64\" Name=\"Points\" NumberOfComponents=\"3\" format=\"binary\"/>\n</PPoints>\n"
           << "<PCells>\n\t<PDataArray type=\"Int32\" Name=\"connectivity\" NumberOfComponents=\"1\" format=\"binary\"/>\n"
           << "\t<PDataArray type=\"Int32\" Name=\"offsets\"      NumberOfComponents=\"1\" format=\"binary\"/>\n"
           << "\t<PDataArray type=\"UInt8\" Name=\"types\"        NumberOfComponents=\"1\" format=\"binary\"/>\n</PCells>\n<PPointData>\n";
        for (const auto& f : fields)
            os << "\t<PDataArray type=\"Float64\" Name=\"" << f.name << "\" NumberOfComponents=\"1\" format=\"binary\" />\n";
        os << "</PPointData>\n<PCellData>\n\t<PDataArray type=\"Int32\" Name=\"attribute\" NumberOfComponents=\"1\" format=\"binary\"/>\n</PCellData>\n"
           << "<Piece Source=\"proc000000.vtu\"/>\n</PUnstructuredGrid>\n</VTKFile>\n";
        if (!os) throw std::runtime_error("cannot write " + cdir + "/data.pvtu");
    }
    // ---- the piece -------------------------------------------------------------------------------------------------------------
    std::vector<int> lex_of_vtk;
    vtk_lagrange_order(dim, p, lex_of_vtk);
    // field at the uniform lattice: 1-D interpolation matrix from the Gauss-Lobatto nodes (identity for p <= 2)
    std::vector<double> uni(D), U, dU;
    for (int a = 0; a < D; ++a) uni[a] = (double)a / p;
    lagrange_tab(s.nodes.empty() ? gauss_lobatto_01(D) : s.nodes, uni, U, dU);      // U[a*D + m] = l_m(a / p)
    const bool identity = p <= 2;
    std::ofstream os(cdir + "/proc000000.vtu");
    os << "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"2.2\" byte_order=\"LittleEndian\">\n<UnstructuredGrid>\n"
       << "<Piece NumberOfPoints=\"" << np << "\" NumberOfCells=\"" << ne << "\">\n";
    {
        std::vector<double> pts((size_t)np * 3, 0.0);
        for (int64_t e = 0; e < ne; ++e) {
            const double* X = &s.elem_corner[(size_t)e * nc * dim];
            for (int l = 0; l < nloc; ++l) {
                const int ijk[3] = {l % D, (l / D) % D, l / (D * D)};
                double* out = &pts[((size_t)e * nloc + l) * 3];
                for (int c = 0; c < nc; ++c) {
                    double w = 1.0;
                    for (int a = 0; a < dim; ++a) w *= ((c >> a) & 1) ? uni[ijk[a]] : 1.0 - uni[ijk[a]];
                    for (int d = 0; d < dim; ++d) out[d] += w * X[c * dim + d];
                }
            }
        }
        os << "<Points>\n<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"binary\">\n";
        block(os, pts);
        os << "</DataArray>\n</Points>\n";
    }
    {
        std::vector<int32_t> conn((size_t)np), offs((size_t)ne);
        std::vector<uint8_t> types((size_t)ne, dim == 3 ? 72 : 70);
        for (int64_t e = 0; e < ne; ++e) {
            for (int j = 0; j < nloc; ++j) conn[(size_t)e * nloc + j] = (int32_t)(e * nloc + lex_of_vtk[j]);
            offs[(size_t)e] = (int32_t)((e + 1) * nloc);
        }
        os << "<Cells>\n<DataArray type=\"Int32\" Name=\"connectivity\" format=\"binary\">\n";
        block(os, conn);
        os << "</DataArray>\n<DataArray type=\"Int32\" Name=\"offsets\" format=\"binary\">\n";
        block(os, offs);
        os << "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"binary\">\n";
        block(os, types);
        os << "</DataArray>\n</Cells>\n";
    }
    {
        std::vector<int32_t> attr((size_t)ne, 1);
        if ((int64_t)elem_attr.size() == ne) for (int64_t e = 0; e < ne; ++e) attr[(size_t)e] = elem_attr[(size_t)e];
        os << "<CellData Scalars=\"attribute\">\n<DataArray type=\"Int32\" Name=\"attribute\" format=\"binary\">\n";
        block(os, attr);
        os << "</DataArray>\n</CellData>\n";
    }
    os << "<PointData >\n";
    std::vector<double> vals((size_t)np), a(nloc), b(nloc);
    for (const auto& f : fields) {
        for (int64_t e = 0; e < ne; ++e) {
            const int* g = &s.gather[(size_t)e * nloc];
            double* out = &vals[(size_t)e * nloc];
            if (identity) { for (int l = 0; l < nloc; ++l) out[l] = f.nodes[g[l]]; continue; }
            for (int l = 0; l < nloc; ++l) a[l] = f.nodes[g[l]];
            // tensor interpolation, one axis at a time
            int stride = 1;
            for (int ax = 0; ax < dim; ++ax, stride *= D) {
                for (int l = 0; l < nloc; ++l) {
                    const int ia = (l / stride) % D, base = l - ia * stride;
                    double sum = 0.0;
                    for (int m = 0; m < D; ++m) sum += U[ia * D + m] * a[base + m * stride];
                    b[l] = sum;
                }
                a.swap(b);
            }
            for (int l = 0; l < nloc; ++l) out[l] = a[l];
        }
        os << "<DataArray type=\"Float64\" Name=\"" << f.name << "\" NumberOfComponents=\"1\" format=\"binary\">\n";
        block(os, vals);
        os << "</DataArray>\n";
    }
    os << "</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n";
    if (!os) throw std::runtime_error("cannot write " + cdir + "/proc000000.vtu");
}

}  // namespace jots
