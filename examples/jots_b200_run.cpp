// Command-line runner over the C ABI: the counterpart of the reference's entry point
// (/root/reference/jots_main.cpp:11-42, `jots -i config.ini`) for the conduction solve alone --
// no output manager, no restart files, no preCICE.  Runs an unmodified JOTS .ini to Max_Timesteps
// on one GPU and prints solver statistics and the temperature range.
//
//   g++ -std=c++17 -Iinclude examples/jots_b200_run.cpp -Ljots_b200 -ljots_b200 -Wl,-rpath,$PWD/jots_b200 -o jots_b200_run
//   ./jots_b200_run -i /path/to/case.ini [-d device] [-n steps]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "jots_b200.h"

int main(int argc, char** argv) {
    const char* ini = nullptr;
    int device = 0, steps = -1;
    for (int i = 1; i < argc; ++i) {
        if ((!std::strcmp(argv[i], "-i") || !std::strcmp(argv[i], "--input")) && i + 1 < argc) ini = argv[++i];
        else if (!std::strcmp(argv[i], "-d") && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "-n") && i + 1 < argc) steps = std::atoi(argv[++i]);
    }
    if (!ini) { std::fprintf(stderr, "usage: %s -i config.ini [-d device] [-n steps]\n", argv[0]); return 2; }
    std::printf("%s\nConfiguration file: %s\n", jots_version(), ini);
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
