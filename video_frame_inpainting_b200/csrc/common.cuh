// Shared host/device helpers for libtai_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <stdio.h>

#include "../../include/tai_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libtai_b200 is written for sm_100a (B200) only"
#endif

namespace tai {

// ---- error plumbing (definitions in capi.cu) ------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);
// which kernel family a separable-convolution launcher chose ("fwd:v3", "bwd_i:v4", "fwd:tiled", ...): read back
// through tai_b200_last_path() by the op sweep
void note_path(const char *path);

// Optional per-kernel timing (CUDA events on the launching stream), switched on by bench.py through
// tai_b200_timing_enable().  `flops` / `bytes` are the ALGORITHMIC work of the launch.
void timing_begin(const char *name, cudaStream_t st, double flops, double bytes);
void timing_end(cudaStream_t st);
struct TimingScope {
    cudaStream_t st;
    TimingScope(const char *name, cudaStream_t s, double flops, double bytes) : st(s) { timing_begin(name, s, flops, bytes); }
    ~TimingScope() { timing_end(st); }
};

inline int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return TAI_ERR_CUDA;
    }
    count_launch();
    return TAI_OK;
}

#define TAI_REQUIRE(cond, code, ...)     \
    do {                                 \
        if (!(cond)) {                   \
            tai::set_error(__VA_ARGS__); \
            return (code);               \
        }                                \
    } while (0)

inline bool fits_int31(long long n) { return n > 0 && n < (1LL << 31); }

// Launch configuration is PER DEVICE (cudaFuncSetAttribute applies to the current device only) and the
// library may be called from several host threads: small per-device tables of atomics, filled on first use
// (the fill is idempotent, so a race only repeats it).
constexpr int kMaxDevices = 64;

inline int current_device()
{
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

inline int sm_count()
{
    static std::atomic<int> cached[kMaxDevices];
    const int dev = current_device();
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

// One table per launcher (a `static KernelConfig cfg;` next to the launch): raises the dynamic shared-memory
// limit of `kern` on the current device once and remembers how many CTAs of it fit on an SM.
struct KernelConfig {
    std::atomic<int> ctas[kMaxDevices];
    // > 0: resident CTAs per SM; -1: the kernel cannot run with `smem` bytes on this device
    template <class K>
    int get(K kern, size_t smem, int nthreads)
    {
        const int dev = current_device();
        int c = ctas[dev].load(std::memory_order_acquire);
        if (c == 0) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                cudaGetLastError();
                c = -1;
            } else {
                int occ = 0;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nthreads, smem);
                c = occ > 0 ? occ : 1;
            }
            ctas[dev].store(c, std::memory_order_release);
        }
        return c;
    }
};

__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Streaming loads for data that is read exactly once (V / H kernel maps): keep them out of L1.
__device__ __forceinline__ float ld_stream(const float *p)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ float4 ld_stream4(const float4 *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

// Packed FP32x2 FMA (Blackwell FFMA2): d = a * b + c on both halves with one issue slot.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    float2 d;
    asm("{\n\t"
        ".reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "mov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t"
        "}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

__device__ __forceinline__ float2 fmul2(float2 a, float2 b)
{
    float2 d;
    asm("{\n\t"
        ".reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t"
        "}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}

}  // namespace tai
