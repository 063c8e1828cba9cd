// Host shim for the hot-path callers of the reference's driver: the .ini key set and defaults of
// Config (/root/reference/src/config_file.cpp:7-172) and the per-step order of JOTSDriver::Run
// (/root/reference/src/jots_driver.cpp:568-725): UpdateAndApplyBCs -> UpdateMatProps -> Iteration.
// Output, restart and preCICE stay with the reference's own driver.
#pragma once
#include "paraview.hpp"
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "operators.hpp"
#include "partition.hpp"

namespace jots {
struct Mesh;
struct Space;

struct Config {
    explicit Config(const std::string& path);
    std::string path, sim_type = "Unsteady", mesh_file = "mesh_name.mesh", time_scheme = "Euler_Implicit";
    std::string solver = "FGMRES", prec = "Chebyshev";
    bool use_restart = false, with_precice = false, using_time_integration = false, using_newton = false;
    int fe_order = 1, serial_refine = 0, parallel_refine = 0, max_timesteps = 100, max_iter = 100, newton_max_iter = 0;
    double initial_temp = 100.0, dt = 0.1, abs_tol = 1e-16, rel_tol = 1e-10, newton_abs_tol = 1e-16, newton_rel_tol = 1e-10;
    // Print_Level lists (config_file.cpp:140-141,153,161-162) and TimeIntegration.Print_Freq (:128)
    std::vector<std::string> ls_print_level{"Errors", "Warnings"}, newton_print_level{"Errors", "Warnings", "Iterations"};
    int time_print_freq = 1;
    // restart files (config_file.cpp:45-46,166-170): VisIt data collection <Restart_Prefix>_<cycle>
    std::string restart_prefix = "restart";
    int restart_cycle = 0, restart_freq = 10, vis_freq = 10;
    std::map<std::string, std::vector<std::string>> mat_props;
    std::map<int, std::vector<std::string>> bcs;
};

Poly parse_property(const std::vector<std::string>& info);
// Factory::CreatePrintLevel (jots_common.inl:57-91): label list -> PRINT_* mask of solvers.hpp; an unknown label is an error
int create_print_level(const std::vector<std::string>& labels);
BCSpec parse_bc(const std::vector<std::string>& info);

class JOTSDriver {
public:
    // mesh == nullptr: read cfg.mesh_file (relative paths are taken as given) and refine
    JOTSDriver(const Config& cfg, int device, const Mesh* mesh = nullptr, struct Comm* comm = nullptr);
    std::unique_ptr<struct Partition> part;   // subdomain (multi-GPU runs)
    void UpdateAndApplyBCs();
    void UpdateMatProps(bool apply_changes);
    void Iteration();
    void Step();               // one pass of the Run() loop body
    int Run(int max_steps);    // max_steps < 0: run to max_timesteps; returns the number of steps done
    void WriteRestartOutput(); // OutputManager::WriteRestartOutput (output_manager.cpp:89-95), restart.cpp
    void WriteVizOutput();     // OutputManager::WriteVizOutput (output_manager.cpp:81-87), paraview.cpp
    std::unique_ptr<ParaViewCollection> viz;   // only when Visualization_Freq != 0
    void collect_nodes(const double* d_vec, std::vector<double>& nodes);   // device vector -> global nodal values on rank 0
    std::unique_ptr<Mesh> out_mesh;      // host copies for restart / ParaView output (only when one of the frequencies != 0)
    std::unique_ptr<Space> out_space;
    Config cfg;
    std::unique_ptr<Ctx> ctx;
    std::unique_ptr<JOTSIterator> it;
    DArr<double> u;
    int it_num = 0, max_timesteps = 0;
    double time = 0.0, dt = 0.0;
    bool initialized_bcs = false;
    bool check_blowup = true;
    bool has_prop[3] = {false, false, false};
};

}  // namespace jots
