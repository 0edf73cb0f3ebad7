// Process-wide state of libmemb: last-error text and the launch counter.
#include "common.cuh"

namespace memb {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
}  // namespace memb

extern "C" const char* memb_last_error(void) { return memb::g_err; }
extern "C" int memb_version(void) { return 100; }
extern "C" int64_t memb_launch_count(void) { return memb::g_launches.load(); }
