// Library-level C ABI: error string, version, launch counter.
#include "common.cuh"
#include "tclight.h"
#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace tcl {
static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace tcl

extern "C" const char* tcl_last_error(void) { return tcl::g_err; }
extern "C" int tcl_version(void) { return 100; }
extern "C" long long tcl_launch_count(void) { return tcl::g_launches.load(); }
extern "C" void tcl_launch_count_reset(void) { tcl::g_launches.store(0); }
