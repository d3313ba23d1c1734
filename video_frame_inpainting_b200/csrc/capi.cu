// Process-wide plumbing of the C ABI: error string, launch counter, version.
#include <atomic>
#include <stdarg.h>

#include "common.cuh"

namespace tai {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace tai

extern "C" int tai_b200_abi_version(void) { return TAI_B200_ABI_VERSION; }
extern "C" const char *tai_b200_last_error(void) { return tai::g_err; }
extern "C" long long tai_b200_launch_count(void) { return tai::g_launches.load(std::memory_order_relaxed); }
