// Restart files in the layout of the reference's VisItDataCollection (OutputManager::WriteRestartOutput,
// /root/reference/src/output_manager.cpp:15-21,89-95; load path src/jots_driver.cpp:214-257):
//
//   <prefix>_<cycle:06d>.mfem_root                 JSON root: cycle, time, mesh path, field "Temperature"
//   <prefix>_<cycle:06d>/mesh.<rank:06d>           MFEM mesh v1.0
//   <prefix>_<cycle:06d>/Temperature.<rank:06d>    GridFunction: "FiniteElementSpace / FiniteElementCollection: H1_<d>D_P<p> /
//                                                  VDim: 1 / Ordering: 0", then one value per DOF, precision 16
//
// The values are written in MFEM's own DOF numbering of the H1 space on that mesh: vertex DOFs by
// vertex id, then (p-1) DOFs per edge -- edges numbered by first appearance in (element, local edge) order, DOFs running from
// the lower to the higher vertex id --, then (p-1)^2 per face -- faces numbered by first appearance in (element, local face)
// order, DOFs in the native face order of the element that introduced the face --, then (p-1)^dim per element.  This restates
// Mesh::GetElementToEdgeTable / GetElementToFaceTable / FiniteElementSpace::BuildDofToArrays of MFEM 4.7, which are not part of
// the reference tree: the numbering is self-consistent (write -> read is exact to the 16 printed digits) and follows MFEM's
// published scheme, but cannot be checked against MFEM here.  The reference writes PARALLEL_FORMAT (one ParMesh piece per
// rank, with communication groups); this writes the single-domain form: on several GPUs rank 0 writes the whole field.
#include "restart.hpp"

#include <sys/stat.h>

#include <algorithm>
#include <array>
#include <cstdio>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>

namespace jots {

namespace {
const int HEX_EDGES[12][2] = {{0, 1}, {1, 2}, {3, 2}, {0, 3}, {4, 5}, {5, 6}, {7, 6}, {4, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
const int HEX_FACES[6][4] = {{3, 2, 1, 0}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};
const int QUAD_EDGES[4][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}};
// lattice coordinates of the MFEM local vertices
const int HEX_VC[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
const int QUAD_VC[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};

std::string dir_of(const std::string& prefix, int cycle) {
    char buf[32];
    std::snprintf(buf, sizeof(buf), "_%06d", cycle);
    return prefix + buf;
}
std::string piece(const std::string& dir, const char* name, int rank) {
    char buf[32];
    std::snprintf(buf, sizeof(buf), ".%06d", rank);
    return dir + "/" + name + buf;
}
}  // namespace

void mfem_dof_numbering(const Mesh& m, int p, const std::vector<int>& gather, int64_t ndof, std::vector<int64_t>& mfem_of_node) {
    const int dim = m.dim, D = p + 1, nv = m.nvpe();
    const int64_t ne = m.n_elem(), nvert = m.n_verts();
    int nloc = 1;
    for (int a = 0; a < dim; ++a) nloc *= D;
    if ((int64_t)gather.size() != ne * nloc) throw std::runtime_error("mfem_dof_numbering: gather table does not match the mesh");
    mfem_of_node.assign((size_t)ndof, -1);
    auto lex = [&](int i, int j, int k) { return i + D * (j + D * k); };
    auto vlex = [&](int lv) {   // lexicographic index of local vertex lv
        return dim == 2 ? lex(QUAD_VC[lv][0] * p, QUAD_VC[lv][1] * p, 0) : lex(HEX_VC[lv][0] * p, HEX_VC[lv][1] * p, HEX_VC[lv][2] * p);
    };
    // ---- vertices
    for (int64_t e = 0; e < ne; ++e)
        for (int lv = 0; lv < nv; ++lv) mfem_of_node[gather[e * nloc + vlex(lv)]] = m.elems[e * nv + lv];
    int64_t next = nvert;
    if (p < 2) return;
    // ---- edges: first appearance in (element, local edge) order; DOFs from the lower to the higher vertex id
    const int nedge = dim == 2 ? 4 : 12;
    std::map<std::pair<int64_t, int64_t>, int64_t> edge_id;
    for (int64_t e = 0; e < ne; ++e)
        for (int le = 0; le < nedge; ++le) {
            const int a = dim == 2 ? QUAD_EDGES[le][0] : HEX_EDGES[le][0], b = dim == 2 ? QUAD_EDGES[le][1] : HEX_EDGES[le][1];
            const int64_t va = m.elems[e * nv + a], vb = m.elems[e * nv + b];
            const auto key = std::make_pair(std::min(va, vb), std::max(va, vb));
            auto it = edge_id.find(key);
            if (it != edge_id.end()) continue;
            const int64_t id = (int64_t)edge_id.size();
            edge_id.emplace(key, id);
            const int la = vlex(a), lb = vlex(b);
            for (int j = 1; j < p; ++j) {
                // j-th interior node from local vertex a towards b
                const int l = la + (lb - la) / p * j;
                const int along = va < vb ? j - 1 : p - 1 - j;
                mfem_of_node[gather[e * nloc + l]] = nvert + id * (p - 1) + along;
            }
        }
    next = nvert + (int64_t)edge_id.size() * (p - 1);
    // ---- faces (3-D): first appearance in (element, local face) order, native order of the introducing element
    if (dim == 3) {
        std::map<std::array<int64_t, 4>, int64_t> face_id;
        const int64_t pf = (int64_t)(p - 1) * (p - 1);
        for (int64_t e = 0; e < ne; ++e)
            for (int lf = 0; lf < 6; ++lf) {
                std::array<int64_t, 4> key;
                for (int v = 0; v < 4; ++v) key[v] = m.elems[e * nv + HEX_FACES[lf][v]];
                std::sort(key.begin(), key.end());
                if (face_id.count(key)) continue;
                const int64_t id = (int64_t)face_id.size();
                face_id.emplace(key, id);
                int o = 0;
                for (int j = 1; j < p; ++j)
                    for (int i = 1; i < p; ++i, ++o) {
                        int l;
                        switch (lf) {   // H1_HexahedronElement: face-interior loops of GetDofMap (mesh.cpp: h1_dof_map)
                            case 0: l = lex(i, p - j, 0); break;
                            case 1: l = lex(i, 0, j); break;
                            case 2: l = lex(p, i, j); break;
                            case 3: l = lex(p - i, p, j); break;
                            case 4: l = lex(0, p - i, j); break;
                            default: l = lex(i, j, p); break;
                        }
                        mfem_of_node[gather[e * nloc + l]] = next + id * pf + o;
                    }
            }
        next += (int64_t)face_id.size() * pf;
    }
    // ---- element interiors
    int64_t pin = 1;
    for (int a = 0; a < dim; ++a) pin *= (p - 1);
    for (int64_t e = 0; e < ne; ++e) {
        int o = 0;
        if (dim == 2) {
            for (int j = 1; j < p; ++j) for (int i = 1; i < p; ++i, ++o) mfem_of_node[gather[e * nloc + lex(i, j, 0)]] = next + e * pin + o;
        } else {
            for (int k = 1; k < p; ++k) for (int j = 1; j < p; ++j) for (int i = 1; i < p; ++i, ++o)
                mfem_of_node[gather[e * nloc + lex(i, j, k)]] = next + e * pin + o;
        }
    }
    next += ne * pin;
    if (next != ndof) throw std::runtime_error("mfem_dof_numbering: the space is not the H1 space of this mesh");
    for (int64_t n = 0; n < ndof; ++n)
        if (mfem_of_node[n] < 0 || mfem_of_node[n] >= ndof) throw std::runtime_error("mfem_dof_numbering: unnumbered node");
}

void write_mfem_mesh(const std::string& path, const Mesh& m) {
    std::ofstream f(path);
    if (!f) throw std::runtime_error("cannot write mesh file: " + path);
    f.precision(16);
    const int nv = m.nvpe(), nf = m.nvpf();
    f << "MFEM mesh v1.0\n\ndimension\n" << m.dim << "\n\nelements\n" << m.n_elem() << "\n";
    for (int64_t e = 0; e < m.n_elem(); ++e) {
        f << (m.elem_attr.empty() ? 1 : m.elem_attr[e]) << " " << (m.dim == 2 ? 3 : 5);
        for (int v = 0; v < nv; ++v) f << " " << m.elems[e * nv + v];
        f << "\n";
    }
    f << "\nboundary\n" << m.n_bdr() << "\n";
    for (int64_t b = 0; b < m.n_bdr(); ++b) {
        f << m.bdr_attr[b] << " " << (m.dim == 2 ? 1 : 3);
        for (int v = 0; v < nf; ++v) f << " " << m.bdr[b * nf + v];
        f << "\n";
    }
    f << "\nvertices\n" << m.n_verts() << "\n" << m.dim << "\n";
    for (int64_t v = 0; v < m.n_verts(); ++v) {
        for (int d = 0; d < m.dim; ++d) f << (d ? " " : "") << m.verts[v * m.dim + d];
        f << "\n";
    }
    if (!f) throw std::runtime_error("error while writing mesh file: " + path);
}

void write_restart(const std::string& prefix, int cycle, double time, const Mesh& mesh, const Space& space, const double* u_nodes) {
    const std::string dir = dir_of(prefix, cycle);
    if (mkdir(dir.c_str(), 0777) != 0) {
        struct stat st;
        if (stat(dir.c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) throw std::runtime_error("cannot create restart directory: " + dir);
    }
    write_mfem_mesh(piece(dir, "mesh", 0), mesh);
    std::vector<int64_t> num;
    mfem_dof_numbering(mesh, space.p, space.gather, space.ndof, num);
    std::vector<double> vals((size_t)space.ndof);
    for (int64_t n = 0; n < space.ndof; ++n) vals[num[n]] = u_nodes[n];
    {
        std::ofstream f(piece(dir, "Temperature", 0));
        if (!f) throw std::runtime_error("cannot write restart field in " + dir);
        f.precision(16);
        f << "FiniteElementSpace\nFiniteElementCollection: H1_" << mesh.dim << "D_P" << space.p << "\nVDim: 1\nOrdering: 0\n\n";
        for (double v : vals) f << v << "\n";
        if (!f) throw std::runtime_error("error while writing the restart field in " + dir);
    }
    // root file (VisItDataCollection::GetVisItRootString)
    const std::string base = dir.substr(dir.find_last_of('/') == std::string::npos ? 0 : dir.find_last_of('/') + 1);
    std::ofstream r(dir + ".mfem_root");
    if (!r) throw std::runtime_error("cannot write restart root file: " + dir + ".mfem_root");
    r.precision(16);
    r << "{\n    \"dsets\": {\n        \"main\": {\n            \"cycle\": " << cycle << ",\n            \"domains\": 1,\n"
      << "            \"fields\": {\n                \"Temperature\": {\n                    \"path\": \"" << base << "/Temperature.%06d\",\n"
      << "                    \"tags\": {\n                        \"assoc\": \"nodes\",\n                        \"comps\": \"1\",\n"
      << "                        \"lod\": \"" << space.p << "\"\n                    }\n                }\n            },\n"
      << "            \"mesh\": {\n                \"format\": \"0\",\n                \"path\": \"" << base << "/mesh.%06d\",\n"
      << "                \"tags\": {\n                    \"max_lods\": \"" << space.p << "\",\n                    \"spatial_dim\": \"" << mesh.dim
      << "\",\n                    \"topo_dim\": \"" << mesh.dim << "\"\n                }\n            },\n"
      << "            \"time\": " << time << ",\n            \"time_step\": 0\n        }\n    }\n}\n";
}

RestartData read_restart(const std::string& prefix, int cycle) {
    const std::string dir = dir_of(prefix, cycle);
    RestartData d;
    {
        std::ifstream r(dir + ".mfem_root");
        if (!r) throw std::runtime_error("Unable to load restart data collection: " + dir + ".mfem_root");
        std::stringstream ss;
        ss << r.rdbuf();
        const std::string s = ss.str();
        auto number_after = [&](const char* key) {
            const size_t k = s.find(std::string("\"") + key + "\"");
            if (k == std::string::npos) throw std::runtime_error(std::string("restart root file has no \"") + key + "\"");
            const size_t c = s.find(':', k);
            return std::stod(s.substr(c + 1));
        };
        d.cycle = (int)number_after("cycle");
        d.time = number_after("time");
        if ((int)number_after("domains") != 1) throw std::runtime_error("restart data collections with several domains are not supported");
    }
    d.mesh = read_mfem_mesh(piece(dir, "mesh", 0));
    std::ifstream f(piece(dir, "Temperature", 0));
    if (!f) throw std::runtime_error("Unable to load restart field 'Temperature' from " + dir);
    std::string ln;
    d.order = -1;
    while (std::getline(f, ln)) {
        const size_t k = ln.find("FiniteElementCollection:");
        if (k != std::string::npos) {
            const size_t pp = ln.rfind("_P");
            if (ln.find("H1_") == std::string::npos || pp == std::string::npos) throw std::runtime_error("restart field is not an H1 grid function: " + ln);
            d.order = std::stoi(ln.substr(pp + 2));
        }
        if (ln.rfind("Ordering:", 0) == 0) break;
    }
    if (d.order < 1) throw std::runtime_error("restart field has no FiniteElementCollection line");
    double v;
    while (f >> v) d.mfem_values.push_back(v);
    return d;
}

void restart_values_to_nodes(const RestartData& d, const Space& space, std::vector<double>& u_nodes) {
    if ((int64_t)d.mfem_values.size() != space.ndof) throw std::runtime_error("restart field size does not match the H1 space of the restart mesh");
    std::vector<int64_t> num;
    mfem_dof_numbering(d.mesh, space.p, space.gather, space.ndof, num);
    u_nodes.resize((size_t)space.ndof);
    for (int64_t n = 0; n < space.ndof; ++n) u_nodes[n] = d.mfem_values[num[n]];
}

}  // namespace jots
