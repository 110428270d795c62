// Library-level state: last-error string, ABI version, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"
#include "grove_b200.h"

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void grove_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void grove_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int grove_abi_version(void) { return GROVE_B200_ABI_VERSION; }
extern "C" const char* grove_last_error(void) { return g_err; }
extern "C" long long grove_launch_count(void) { return g_launches.load(); }
extern "C" void grove_reset_launch_count(void) { g_launches.store(0); }
extern "C" void grove_add_launch_count(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
