// Library-wide state of libnbe_b200.so: error string, launch counter, ABI version.
#include "common.cuh"
#include <cstdlib>

namespace nbe {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
bool pdl_enabled() {
    static const bool on = getenv("NBE_NO_PDL") == nullptr;
    return on;
}
}  // namespace nbe

extern "C" int nbe_abi_version(void) { return NBE_ABI_VERSION; }
extern "C" const char* nbe_last_error(void) { return nbe::g_err; }
extern "C" int64_t nbe_launch_count(void) { return nbe::g_launches.load(std::memory_order_relaxed); }
