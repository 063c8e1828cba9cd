// Single-GPU build of the interface exchange: contexts never carry a Halo, so these are unreachable.
#include "halo.hpp"

#include <stdexcept>

#include "ctx.hpp"

namespace jots {
struct Halo {};
void halo_sum_exchange(Ctx&, double*) { throw std::runtime_error("multi-GPU interface exchange is not built into this library"); }
void halo_allreduce(Ctx&, double*, int) { throw std::runtime_error("multi-GPU all-reduce is not built into this library"); }
const int64_t* halo_global_index(Ctx&) { return nullptr; }
}  // namespace jots
