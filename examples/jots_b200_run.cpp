// Command-line runner over the C ABI: the counterpart of the reference's entry point
// (/root/reference/jots_main.cpp:11-42, `jots -i config.ini`) for the conduction solve alone --
// no output manager, no restart files, no preCICE.  Runs an unmodified JOTS .ini to Max_Timesteps
// on one GPU and prints solver statistics and the temperature range.
//
//   g++ -std=c++17 -Iinclude examples/jots_b200_run.cpp -Ljots_b200 -ljots_b200 -Wl,-rpath,$PWD/jots_b200 -pthread -o jots_b200_run
//   ./jots_b200_run -i /path/to/case.ini [-d device] [-n steps] [-g gpus]
//
// -g N (the counterpart of `mpirun -np N jots`): N host threads, one subdomain and one NCCL rank per GPU.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "jots_b200.h"

struct RankResult { int rc = 0; std::string err; int done = 0; double tmin = 0, tmax = 0, time = 0; long long n_owned = 0, krylov = 0, newton = 0; };

static void run_rank(const char* ini, int rank, int nranks, const char* uid, int steps, RankResult* out) {
    jots_comm_t comm = jots_comm_create(uid, rank, nranks, rank);
    if (!comm) { out->rc = 1; out->err = jots_last_error(); return; }
    jots_driver_t drv = jots_driver_create_dist(ini, nullptr, rank, comm);
    if (!drv) { out->rc = 1; out->err = jots_last_error(); jots_comm_destroy(comm); return; }
    if (jots_driver_run(drv, steps, &out->done)) { out->rc = 1; out->err = jots_last_error(); }
    else {
        const int64_t n = jots_driver_size(drv);
        std::vector<double> u(n);
        std::vector<int64_t> gid(n);
        int64_t n_owned = 0, st[12];
        double norms[2];
        jots_driver_get_u(drv, u.data(), JOTS_HOST);
        jots_driver_copy_global_ids(drv, gid.data(), &n_owned);
        jots_ctx_stats(jots_driver_ctx(drv), st, norms);
        out->n_owned = n_owned; out->krylov = st[2]; out->newton = st[3]; out->time = jots_driver_time(drv);
        out->tmin = *std::min_element(u.begin(), u.begin() + n_owned);
        out->tmax = *std::max_element(u.begin(), u.begin() + n_owned);
    }
    jots_driver_destroy(drv);
    jots_comm_destroy(comm);
}

int main(int argc, char** argv) {
    const char* ini = nullptr;
    int device = 0, steps = -1, gpus = 1;
    for (int i = 1; i < argc; ++i) {
        if ((!std::strcmp(argv[i], "-i") || !std::strcmp(argv[i], "--input")) && i + 1 < argc) ini = argv[++i];
        else if (!std::strcmp(argv[i], "-d") && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "-n") && i + 1 < argc) steps = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "-g") && i + 1 < argc) gpus = std::atoi(argv[++i]);
    }
    if (!ini) { std::fprintf(stderr, "usage: %s -i config.ini [-d device] [-n steps] [-g gpus]\n", argv[0]); return 2; }
    std::printf("%s\nConfiguration file: %s\n", jots_version(), ini);
    if (gpus > 1) {
        if (gpus > jots_device_count()) { std::fprintf(stderr, "error: %d GPUs requested, %d available\n", gpus, jots_device_count()); return 1; }
        char uid[128];
        if (jots_comm_unique_id(uid)) { std::fprintf(stderr, "error: %s\n", jots_last_error()); return 1; }
        std::vector<RankResult> res(gpus);
        std::vector<std::thread> th;
        for (int r = 0; r < gpus; ++r) th.emplace_back(run_rank, ini, r, gpus, uid, steps, &res[r]);
        for (auto& t : th) t.join();
        long long ndof = 0;
        double tmin = 1e300, tmax = -1e300;
        for (int r = 0; r < gpus; ++r) {
            if (res[r].rc) { std::fprintf(stderr, "error (rank %d): %s\n", r, res[r].err.c_str()); return 1; }
            ndof += res[r].n_owned; tmin = std::min(tmin, res[r].tmin); tmax = std::max(tmax, res[r].tmax);
        }
        std::printf("GPUs: %d   Steps: %d   Time: %.6g   DOFs: %lld\n", gpus, res[0].done, res[0].time, ndof);
        std::printf("Newton iterations: %lld   Krylov iterations: %lld\n", res[0].newton, res[0].krylov);
        std::printf("Temperature: min %.10g  max %.10g\n", tmin, tmax);
        return 0;
    }
    jots_driver_t drv = jots_driver_create(ini, device);
    if (!drv) { std::fprintf(stderr, "error: %s\n", jots_last_error()); return 1; }   // MFEM_ABORT in the reference
    int done = 0;
    if (jots_driver_run(drv, steps, &done)) { std::fprintf(stderr, "error: %s\n", jots_last_error()); jots_driver_destroy(drv); return 1; }
    const int64_t n = jots_driver_size(drv);
    std::vector<double> u(n);
    if (jots_driver_get_u(drv, u.data(), JOTS_HOST)) { std::fprintf(stderr, "error: %s\n", jots_last_error()); return 1; }
    int64_t st[12];
    double norms[2];
    jots_ctx_stats(jots_driver_ctx(drv), st, norms);
    std::printf("Steps: %d   Time: %.6g   DOFs: %lld\n", done, jots_driver_time(drv), (long long)n);
    std::printf("Newton iterations: %lld   Krylov iterations: %lld   operator applies: %lld   kernel launches: %lld\n",
                (long long)st[3], (long long)st[2], (long long)st[0], (long long)st[6]);
    std::printf("Temperature: min %.10g  max %.10g\n", *std::min_element(u.begin(), u.end()), *std::max_element(u.begin(), u.end()));
    jots_driver_destroy(drv);
    return 0;
}
