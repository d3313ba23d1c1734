// Process-wide plumbing of the C ABI: error string, launch counter, version.
#include <atomic>
#include <map>
#include <mutex>
#include <stdarg.h>
#include <string.h>
#include <string>
#include <vector>

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

static thread_local const char *g_path = "";
void note_path(const char *path) { g_path = path; }

// ---- per-kernel timing ---------------------------------------------------------------------------
struct TimingRecord {
    std::string name;
    cudaEvent_t e0, e1;
    double flops, bytes;
};
static std::atomic<int> g_timing_on{0};
static std::mutex g_timing_mu;
static std::vector<TimingRecord> g_timing;
static thread_local int g_timing_open = -1;

void timing_begin(const char *name, cudaStream_t st, double flops, double bytes)
{
    if (!g_timing_on.load(std::memory_order_relaxed)) return;
    TimingRecord r{name, nullptr, nullptr, flops, bytes};
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, st);
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_timing.push_back(r);
    g_timing_open = (int)g_timing.size() - 1;
}

void timing_end(cudaStream_t st)
{
    if (g_timing_open < 0) return;
    std::lock_guard<std::mutex> lk(g_timing_mu);
    if (g_timing_open < (int)g_timing.size()) cudaEventRecord(g_timing[g_timing_open].e1, st);
    g_timing_open = -1;
}

}  // namespace tai

extern "C" int tai_b200_abi_version(void) { return TAI_B200_ABI_VERSION; }
extern "C" const char *tai_b200_last_error(void) { return tai::g_err; }
extern "C" long long tai_b200_launch_count(void) { return tai::g_launches.load(std::memory_order_relaxed); }
extern "C" const char *tai_b200_last_path(void) { return tai::g_path; }

extern "C" int tai_b200_timing_enable(int on)
{
    std::lock_guard<std::mutex> lk(tai::g_timing_mu);
    for (auto &r : tai::g_timing) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    tai::g_timing.clear();
    tai::g_timing_on.store(on ? 1 : 0);
    return TAI_OK;
}

// Waits for the recorded events and writes a JSON array
// [{"name":..,"launches":n,"ms":total,"flops":total,"bytes":total}, ...] into buf.
extern "C" int tai_b200_timing_report(char *buf, int buflen)
{
    TAI_REQUIRE(buf && buflen > 2, TAI_ERR_INVALID_ARGUMENT, "tai_b200_timing_report: bad buffer");
    struct Agg { long n = 0; double ms = 0, flops = 0, bytes = 0; };
    std::map<std::string, Agg> agg;
    {
        std::lock_guard<std::mutex> lk(tai::g_timing_mu);
        for (auto &r : tai::g_timing) {
            float ms = 0.f;
            if (cudaEventSynchronize(r.e1) != cudaSuccess || cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) {
                cudaGetLastError();
                continue;
            }
            Agg &a = agg[r.name];
            a.n += 1; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
        }
    }
    std::string out = "[";
    bool first = true;
    for (auto &kv : agg) {
        char line[512];
        snprintf(line, sizeof(line), "%s{\"name\":\"%s\",\"launches\":%ld,\"ms\":%.6f,\"flops\":%.6e,\"bytes\":%.6e}",
                 first ? "" : ",", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.flops, kv.second.bytes);
        out += line;
        first = false;
    }
    out += "]";
    TAI_REQUIRE((int)out.size() + 1 <= buflen, TAI_ERR_INVALID_ARGUMENT, "tai_b200_timing_report: buffer too small");
    memcpy(buf, out.c_str(), out.size() + 1);
    return TAI_OK;
}
