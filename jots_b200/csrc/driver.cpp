#include "driver.hpp"

#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include <cstdio>

#include "halo.hpp"
#include "restart.hpp"

namespace jots {

// JOTS_QUIET=1 silences the driver's and the solvers' console output (the config keys Print_Level / Print_Freq still parse)
bool quiet() {
    static const bool q = [] { const char* e = std::getenv("JOTS_QUIET"); return e && e[0] == '1'; }();
    return q;
}

static std::string trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n");
    if (a == std::string::npos) return "";
    size_t b = s.find_last_not_of(" \t\r\n");
    return s.substr(a, b - a + 1);
}
static std::vector<std::string> split_list(const std::string& s) {
    std::vector<std::string> out;
    std::stringstream ss(s);
    std::string t;
    while (std::getline(ss, t, ',')) out.push_back(trim(t));
    return out;
}
typedef std::map<std::string, std::vector<std::pair<std::string, std::string>>> Ini;

static Ini read_ini(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open config file: " + path);
    Ini ini;
    std::string ln, sec;
    while (std::getline(f, ln)) {
        ln = trim(ln);
        if (ln.empty() || ln[0] == ';' || ln[0] == '#') continue;
        if (ln[0] == '[') { sec = trim(ln.substr(1, ln.find(']') - 1)); ini[sec]; continue; }
        size_t eq = ln.find('=');
        if (eq == std::string::npos) continue;
        ini[sec].emplace_back(trim(ln.substr(0, eq)), trim(ln.substr(eq + 1)));
    }
    return ini;
}
static const std::string* find_key(const Ini& ini, const std::string& sec, const std::string& key) {
    auto it = ini.find(sec);
    if (it == ini.end()) return nullptr;
    for (auto& kv : it->second) if (kv.first == key) return &kv.second;
    return nullptr;
}
static std::string get_str(const Ini& ini, const std::string& sec, const std::string& key, const std::string& def) {
    const std::string* v = find_key(ini, sec, key);
    return v ? *v : def;
}
// boost ptree get(key, default): falls back to the default when the text does not parse
// completely as the requested type (e.g. Max_Timesteps=5e6 -> 100)
static int get_int(const Ini& ini, const std::string& sec, const std::string& key, int def) {
    const std::string* v = find_key(ini, sec, key);
    if (!v) return def;
    char* end = nullptr;
    long r = std::strtol(v->c_str(), &end, 10);
    if (end == v->c_str() || *end != '\0') return def;
    return (int)r;
}
static double get_dbl(const Ini& ini, const std::string& sec, const std::string& key, double def) {
    const std::string* v = find_key(ini, sec, key);
    if (!v) return def;
    char* end = nullptr;
    double r = std::strtod(v->c_str(), &end);
    if (end == v->c_str() || *end != '\0') return def;
    return r;
}

Config::Config(const std::string& p) : path(p) {
    Ini ini = read_ini(p);
    const char* FE = "FiniteElementSetup";
    sim_type = get_str(ini, FE, "Simulation_Type", "Unsteady");
    use_restart = get_str(ini, FE, "Use_Restart", "No") == "Yes";
    mesh_file = get_str(ini, FE, "Mesh_File", "mesh_name.mesh");
    fe_order = get_int(ini, FE, "FE_Order", 1);
    serial_refine = get_int(ini, FE, "Serial_Refine", 0);
    parallel_refine = get_int(ini, FE, "Parallel_Refine", 0);
    initial_temp = get_dbl(ini, FE, "Initial_Temperature", 100.0);
    restart_prefix = get_str(ini, FE, "Restart_Prefix", "restart");
    restart_cycle = get_int(ini, FE, "Restart_Cycle", 0);
    if (!ini.count("MaterialProperties")) throw std::runtime_error("config file has no [MaterialProperties] section");
    for (auto& kv : ini["MaterialProperties"]) mat_props[kv.first] = split_list(kv.second);
    with_precice = ini.count("preCICE") > 0;
    if (!ini.count("BoundaryConditions")) throw std::runtime_error("config file has no [BoundaryConditions] section");
    for (auto& kv : ini["BoundaryConditions"]) {
        const size_t us = kv.first.find_last_of('_');
        bcs[std::stoi(kv.first.substr(us + 1))] = split_list(kv.second);
    }
    using_time_integration = ini.count("TimeIntegration") > 0;
    if (using_time_integration) {
        time_scheme = get_str(ini, "TimeIntegration", "Time_Scheme", "Euler_Implicit");
        dt = get_dbl(ini, "TimeIntegration", "Delta_Time", 0.1);
        max_timesteps = get_int(ini, "TimeIntegration", "Max_Timesteps", 100);
        time_print_freq = get_int(ini, "TimeIntegration", "Print_Freq", 1);
    }
    const char* LS = "LinearSolverSettings";
    solver = get_str(ini, LS, "Solver", "FGMRES");
    prec = get_str(ini, LS, "Preconditioner", "Chebyshev");
    abs_tol = get_dbl(ini, LS, "Absolute_Tolerance", 1e-16);
    rel_tol = get_dbl(ini, LS, "Relative_Tolerance", 1e-10);
    max_iter = get_int(ini, LS, "Max_Iterations", 100);
    restart_freq = get_int(ini, "Output", "Restart_Freq", 10);
    vis_freq = get_int(ini, "Output", "Visualization_Freq", 10);
    ls_print_level = split_list(get_str(ini, LS, "Print_Level", "Errors, Warnings"));
    using_newton = ini.count("NewtonSolverSettings") > 0;
    if (using_newton) {
        const char* NS = "NewtonSolverSettings";
        newton_max_iter = get_int(ini, NS, "Max_Iterations", 100);
        newton_abs_tol = get_dbl(ini, NS, "Absolute_Tolerance", 1e-16);
        newton_rel_tol = get_dbl(ini, NS, "Relative_Tolerance", 1e-10);
        newton_print_level = split_list(get_str(ini, NS, "Print_Level", "Errors, Warnings, Iterations"));
    }
}

int create_print_level(const std::vector<std::string>& labels) {
    // IterativeSolver::PrintLevel: every label switches its flag on, "None" clears, "All" sets everything
    int m = 0;
    for (const std::string& l : labels) {
        if (l == "None") m = 0;
        else if (l == "Warnings") m |= PRINT_WARNINGS;
        else if (l == "Errors") m |= PRINT_ERRORS;
        else if (l == "Iterations") m |= PRINT_ITERATIONS;
        else if (l == "FirstAndLast") m |= PRINT_FIRST_AND_LAST;
        else if (l == "Summary") m |= PRINT_SUMMARY;
        else if (l == "All") m = PRINT_ERRORS | PRINT_WARNINGS | PRINT_ITERATIONS | PRINT_SUMMARY | PRINT_FIRST_AND_LAST;
        else throw std::runtime_error("Unknown print level: '" + l + "'");     // Print_Level_Map.at() throws in the reference
    }
    return m;
}

Poly parse_property(const std::vector<std::string>& info) {
    Poly p;
    p.n = 0;
    for (double& v : p.c) v = 0.0;
    if (info.empty()) throw std::runtime_error("empty material property specification");
    if (info[0] == "Uniform") {
        if (info.size() < 2) throw std::runtime_error("Uniform property needs a value");
        p.n = 1; p.c[0] = std::stod(info[1]);
    } else if (info[0] == "Polynomial") {
        const int n = (int)info.size() - 1;
        if (n < 2) throw std::runtime_error("Polynomial property needs >= 2 coefficients (material_property.cpp:70)");
        if (n > MAX_POLY) throw std::runtime_error("Polynomial property with more than 8 coefficients is not supported");
        p.n = n;
        for (int i = 0; i < n; ++i) p.c[i] = std::stod(info[1 + i]);
    } else throw std::runtime_error("Unknown/Invalid material model specified");
    return p;
}

BCSpec parse_bc(const std::vector<std::string>& info) {
    BCSpec b;
    if (info.empty()) throw std::runtime_error("empty boundary condition specification");
    const std::string& k = info[0];
    if (k == "Isothermal" || k == "HeatFlux") {
        if (info.size() < 2) throw std::runtime_error("boundary condition needs a value");
        b.kind = k == "Isothermal" ? BC_ISOTHERMAL : BC_HEATFLUX;
        b.v[0] = std::stod(info[1]);
    } else if (k == "Sinusoidal_Isothermal" || k == "Sinusoidal_HeatFlux") {
        if (info.size() < 5) throw std::runtime_error("sinusoidal boundary condition needs 4 parameters");
        b.kind = k == "Sinusoidal_Isothermal" ? BC_SIN_ISOTHERMAL : BC_SIN_HEATFLUX;
        for (int i = 0; i < 4; ++i) b.v[i] = std::stod(info[1 + i]);
    } else if (k == "preCICE_Isothermal" || k == "preCICE_HeatFlux") {
        // kind, coupling-mesh name, default value (jots_driver.cpp:451-462); the adapter moves the data through
        // jots_ctx_precice_bc_set_read_data / _write_data
        if (info.size() < 3) throw std::runtime_error("preCICE boundary condition needs a mesh name and a default value");
        b.kind = k == "preCICE_Isothermal" ? BC_PRECICE_ISOTHERMAL : BC_PRECICE_HEATFLUX;
        b.v[0] = std::stod(info[2]);
    } else throw std::runtime_error("Invalid/Unknown boundary condition specified: '" + k + "'");
    return b;
}

static int map_label(const std::string& s, const std::vector<std::pair<std::string, int>>& table, const char* what) {
    for (auto& kv : table) if (kv.first == s) return kv.second;
    throw std::runtime_error(std::string("unknown ") + what + ": '" + s + "'");
}

JOTSDriver::JOTSDriver(const Config& cfg_, int device, const Mesh* mesh_in, Comm* comm) : cfg(cfg_) {
    if (cfg.with_precice)
        throw std::runtime_error("preCICE configurations are handled by the reference's own driver (out of scope)");
    Mesh mesh;
    RestartData restart;
    int order = cfg.fe_order;
    if (cfg.use_restart) {
        // jots_driver.cpp:214-257: mesh, FE space and Temperature come from the VisIt data collection <prefix>_<cycle>,
        // not from Mesh_File / FE_Order / Initial_Temperature
        restart = read_restart(cfg.restart_prefix, cfg.restart_cycle);
        mesh = restart.mesh;
        order = restart.order;
    } else if (mesh_in) mesh = *mesh_in;
    else {
        mesh = read_mfem_mesh(cfg.mesh_file);
        for (int i = 0; i < cfg.serial_refine + cfg.parallel_refine; ++i) mesh = uniform_refine(mesh);
    }
    Space space = make_space(mesh, order);
    if (cfg.restart_freq != 0 || cfg.vis_freq != 0) {
        // the output writers need the mesh and the (global) space on the host: kept only when something is written
        out_mesh.reset(new Mesh(mesh));
        out_space.reset(new Space(space));
    }
    if (cfg.vis_freq != 0) {
        viz.reset(new ParaViewCollection());
        viz->restart_mode = cfg.use_restart;     // paraview_dc->UseRestartMode(UsesRestart()) (output_manager.cpp:25)
        if (const char* d = std::getenv("JOTS_PARAVIEW_DIR")) viz->dir = d;
    }
    if (comm && comm_size(comm) > 1) {
        // ParMesh(comm, mesh): one subdomain per rank (jots_driver.cpp:178)
        std::vector<int> elem_rank;
        partition_elements(mesh, comm_size(comm), nullptr, elem_rank);
        part.reset(new Partition(make_partition(space, elem_rank, comm_size(comm), comm_rank(comm))));
        ctx.reset(new Ctx(*part, device, comm));
    } else ctx.reset(new Ctx(space, device));
    // material properties
    Poly one; one.n = 1; for (double& v : one.c) v = 0.0; one.c[0] = 1.0;
    Poly props[3] = {one, one, one};
    const char* keys[3] = {"Density_rho", "Specific_Heat_C", "Thermal_Conductivity_k"};
    for (int i = 0; i < 3; ++i) {
        auto f = cfg.mat_props.find(keys[i]);
        if (f != cfg.mat_props.end()) { props[i] = parse_property(f->second); has_prop[i] = true; }
    }
    ctx->set_materials(props[0], props[1], props[2]);
    // boundary conditions: one per mesh boundary attribute, index i <-> attribute i+1
    const std::vector<int>& attrs = ctx->space.attributes;
    if (attrs.size() != cfg.bcs.size()) throw std::runtime_error("Input file BC count and mesh BC counts do not match.");
    std::vector<BCSpec> bcs;
    for (int a : attrs) {
        auto f = cfg.bcs.find(a);
        if (f == cfg.bcs.end()) throw std::runtime_error("No matching boundary attribute in config file for attribute " + std::to_string(a));
        bcs.push_back(parse_bc(f->second));
    }
    if (cfg.using_time_integration) { it_num = 0; max_timesteps = cfg.max_timesteps; dt = cfg.dt; }
    else { it_num = -2; max_timesteps = -1; dt = 0.0; }
    time = 0.0;
    if (cfg.use_restart) { it_num = restart.cycle; time = restart.time; }     // jots_driver.cpp:246-247
    for (auto& b : bcs) b.update(time);
    ctx->set_bcs(bcs);
    u.alloc(ctx->n);
    if (cfg.use_restart) {
        std::vector<double> nodes, local((size_t)ctx->n);
        restart_values_to_nodes(restart, space, nodes);
        for (int64_t i = 0; i < ctx->n; ++i) local[i] = nodes[part ? part->global_id[i] : i];
        JOTS_CUDA(cudaMemcpyAsync(u.p, local.data(), (size_t)ctx->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        ctx->sync();
    } else v_fill(ctx->stream, ctx->n, cfg.initial_temp, u.p);
    UpdateMatProps(false);
    SolverSpec ls;
    ls.solver = map_label(cfg.solver, {{"CG", SOLVER_CG}, {"GMRES", SOLVER_GMRES}, {"FGMRES", SOLVER_FGMRES}}, "solver");
    ls.prec = map_label(cfg.prec, {{"Jacobi", PREC_JACOBI}, {"Chebyshev", PREC_CHEBYSHEV}}, "preconditioner");
    ls.rel_tol = cfg.rel_tol; ls.abs_tol = cfg.abs_tol; ls.max_iter = cfg.max_iter;
    ls.print = create_print_level(cfg.ls_print_level);
    NewtonSpec nw;
    nw.rel_tol = cfg.newton_rel_tol; nw.abs_tol = cfg.newton_abs_tol; nw.max_iter = cfg.newton_max_iter;
    nw.print = create_print_level(cfg.newton_print_level);
    const int scheme = map_label(cfg.time_scheme, {{"Euler_Implicit", TS_EULER_IMPLICIT}, {"Euler_Explicit", TS_EULER_EXPLICIT}, {"RK4", TS_RK4}}, "time scheme");
    if (cfg.sim_type == "Linearized_Unsteady") {
        if (!(has_prop[0] && has_prop[1] && has_prop[2])) throw std::runtime_error("Linearized_Unsteady needs Density_rho, Specific_Heat_C and Thermal_Conductivity_k");
        it.reset(new LinearConductionOperator(ctx.get(), ls, dt, scheme));
    } else if (cfg.sim_type == "Nonlinear_Unsteady") {
        if (!(has_prop[0] && has_prop[1] && has_prop[2])) throw std::runtime_error("Nonlinear_Unsteady needs Density_rho, Specific_Heat_C and Thermal_Conductivity_k");
        it.reset(new NonlinearConductionOperator(ctx.get(), ls, nw, dt, scheme));
    } else if (cfg.sim_type == "Steady") {
        if (!has_prop[2]) throw std::runtime_error("Steady needs Thermal_Conductivity_k");
        it.reset(new SteadyConductionOperator(ctx.get(), ls, nw));
    } else throw std::runtime_error("unknown Simulation_Type: '" + cfg.sim_type + "'");
}

void JOTSDriver::UpdateMatProps(bool apply_changes) {
    // jots_driver.cpp:656-671; MATERIAL_PROPERTY enum order
    const Poly* props[3] = {&ctx->mat.rho, &ctx->mat.C, &ctx->mat.k};
    bool any = false;
    for (int i = 0; i < 3; ++i) any = any || (has_prop[i] && props[i]->n > 1);
    if (!any) { if (!ctx->mat_state.p) ctx->set_mat_state(u.p); return; }
    ctx->set_mat_state(u.p);
    if (apply_changes)
        for (int i = 0; i < 3; ++i)
            if (has_prop[i] && props[i]->n > 1) it->ProcessMatPropUpdate(i);
}

void JOTSDriver::UpdateAndApplyBCs() {
    // jots_driver.cpp:673-716: values at t_n; later BCs overwrite earlier ones on shared DOFs
    bool n_changed = false;
    for (size_t i = 0; i < ctx->bcs.size(); ++i) {
        BCSpec& bc = ctx->bcs[i];
        if (!bc.constant() || !initialized_bcs) {
            bc.update(time);
            if (bc.essential()) ctx->apply_dirichlet((int)i, u.p);
            else n_changed = true;
        }
    }
    initialized_bcs = true;
    if (n_changed) it->UpdateNeumann();
}

void JOTSDriver::Iteration() {
    it->time = time;
    it->Iterate(u.p);
    time = it->time;
    ++it_num;
}

void JOTSDriver::Step() {
    NvtxRange range("jots:step");
    UpdateAndApplyBCs();
    UpdateMatProps(true);
    Iteration();
    // PostprocessIteration (jots_driver.cpp:727-743): step line every Print_Freq steps on rank 0, then the blow-up check.  The
    // reference tests the rank-local maximum and MFEM_ABORT takes every rank down through MPI_Abort; here the maximum is
    // all-reduced so that every rank throws together instead of leaving the others in the next collective.
    if (cfg.using_time_integration && cfg.time_print_freq > 0 && it_num % cfg.time_print_freq == 0 && ctx->rank() == 0 && !quiet())
        std::printf("Step #%10i || Time: %1.6e out of %-1.6e || dt: %1.6e \n", it_num, time, dt * max_timesteps, dt);
    if (check_blowup && ctx->max_global(u.p) > 1e10) throw std::runtime_error("JOTS has blown up");
    // output at the input frequencies (jots_driver.cpp:773-800): properties refreshed at the new field, ParaView, then restart
    if (cfg.using_time_integration) {
        const bool viz_out = cfg.vis_freq != 0 && it_num % cfg.vis_freq == 0;
        const bool res_out = cfg.restart_freq != 0 && it_num % cfg.restart_freq == 0;
        if (viz_out || res_out) {
            UpdateMatProps(false);
            if (viz_out) WriteVizOutput();
            if (res_out) WriteRestartOutput();
        }
    }
}

int JOTSDriver::Run(int max_steps) {
    int done = 0;
    // initial condition to ParaView (jots_driver.cpp:571-573).  Deviation: the reference writes it whenever the run integrates in
    // time; here only when Visualization_Freq != 0, so that a run that asks for no output writes no files
    if (it_num == 0 && cfg.using_time_integration && viz) WriteVizOutput();
    while (it_num < max_timesteps) {
        Step();
        ++done;
        if (max_steps >= 0 && done >= max_steps) break;
    }
    // steady run: output once at the end when the frequencies are non-zero (jots_driver.cpp:750-772)
    if (!cfg.using_time_integration && done > 0 && (viz || cfg.restart_freq != 0)) {
        UpdateMatProps(false);
        if (viz) WriteVizOutput();
        if (cfg.restart_freq != 0) WriteRestartOutput();
    }
    ctx->sync();
    return done;
}

void JOTSDriver::collect_nodes(const double* d_vec, std::vector<double>& nodes) {
    if (part) {
        // rank 0 collects the owned values of every rank (single-domain output files); collective
        std::vector<int64_t> gid;
        std::vector<double> val;
        halo_gather_owned(*ctx, d_vec, gid, val);
        nodes.clear();
        if (ctx->rank() != 0) return;
        nodes.assign((size_t)out_space->ndof, 0.0);
        for (size_t i = 0; i < gid.size(); ++i) nodes[(size_t)gid[i]] = val[i];
    } else {
        nodes.resize((size_t)ctx->n);
        JOTS_CUDA(cudaMemcpyAsync(nodes.data(), d_vec, (size_t)ctx->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->sync();
    }
}

void JOTSDriver::WriteRestartOutput() {
    if (!out_mesh) throw std::runtime_error("restart output was not set up (Restart_Freq was 0 when the driver was built)");
    std::vector<double> nodes;
    collect_nodes(u.p, nodes);
    if (ctx->rank() != 0) return;
    if (!quiet()) std::printf("Saving Restart File...\n");
    write_restart(cfg.restart_prefix, it_num, time, *out_mesh, *out_space, nodes.data());
}

// Fields of the reference's collection: "Rank" (output_manager.cpp:33), "Temperature" (jots_driver.cpp:288) and one field per
// material property of the input file under its own key (jots_driver.cpp:343), each the nodal projection of the property
// coefficient at the last material state (OutputManager::UpdateGridFunctions)
void JOTSDriver::WriteVizOutput() {
    if (!viz || !out_space) throw std::runtime_error("ParaView output was not set up (Visualization_Freq was 0 when the driver was built)");
    std::vector<std::pair<std::string, std::vector<double>>> host;
    DArr<double> tmp;
    tmp.alloc((size_t)ctx->n);
    auto add = [&](const std::string& name, const double* d_vec) {
        host.emplace_back(name, std::vector<double>());
        collect_nodes(d_vec, host.back().second);
    };
    v_fill(ctx->stream, ctx->n, (double)ctx->rank(), tmp.p);
    add("Rank", tmp.p);
    add("Temperature", u.p);
    const char* keys[3] = {"Density_rho", "Specific_Heat_C", "Thermal_Conductivity_k"};
    const Poly* props[3] = {&ctx->mat.rho, &ctx->mat.C, &ctx->mat.k};
    for (int i = 0; i < 3; ++i)
        if (has_prop[i]) {
            v_poly_field(ctx->stream, ctx->n, *props[i], 0, props[i]->n > 1 ? ctx->mat_state.p : nullptr, tmp.p);
            ++ctx->stats.kernel_launches;
            add(keys[i], tmp.p);
        }
    if (ctx->rank() != 0) return;
    if (!quiet()) std::printf("Saving Paraview Data...\n");
    std::vector<VizField> fields;
    for (auto& h : host) fields.push_back({h.first, h.second.data()});
    viz->save(it_num, time, *out_space, out_mesh->elem_attr, fields);
}

}  // namespace jots
