// Microbenchmark (development aid): FFMA / FFMA2 issue rate vs number of fresh register operands.
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"
namespace tai { void set_error(const char *, ...) {} void count_launch(int) {} }
using namespace tai;

// MODE 0: a = fma(a, s, t)            (1 fresh register operand per FMA)
// MODE 1: a[k] = fma(b[k], s, a[k])   (2 fresh, 8 b registers)
// MODE 2: a[k] = fma(b[k][j], s[j], a[k]) j<12  (2 fresh, 96 b registers; s[j] reused across k)
template <int MODE, bool PACKED>
__global__ void __launch_bounds__(128, 3) k(const float *src, float *out, int iters)
{
    float b[8][12];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int j = 0; j < 12; ++j) b[r][j] = src[(r * 12 + j) * 128 + threadIdx.x];
    float s[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) s[j] = src[j];
    float a[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    float2 a2[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) a2[r] = make_float2(r, r + 0.5f);
    const float t = src[5];
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (!PACKED) {
            if (MODE == 0) {
#pragma unroll
                for (int j = 0; j < 12; ++j)
#pragma unroll
                    for (int r = 0; r < 8; ++r) a[r] = fmaf(a[r], s[0], t);
            } else if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < 12; ++j)
#pragma unroll
                    for (int r = 0; r < 8; ++r) a[r] = fmaf(b[r][0], s[0], a[r]);
            } else {
#pragma unroll
                for (int j = 0; j < 12; ++j)
#pragma unroll
                    for (int r = 0; r < 8; ++r) a[r] = fmaf(b[r][j], s[j], a[r]);
            }
        } else {
            if (MODE == 0) {
#pragma unroll
                for (int j = 0; j < 6; ++j)
#pragma unroll
                    for (int r = 0; r < 8; ++r) a2[r] = ffma2(a2[r], make_float2(s[0], s[1]), make_float2(t, t));
            } else if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < 6; ++j)
#pragma unroll
                    for (int r = 0; r < 8; ++r) a2[r] = ffma2(make_float2(b[r][0], b[r][1]), make_float2(s[0], s[1]), a2[r]);
            } else {
#pragma unroll
                for (int j = 0; j < 6; ++j)
#pragma unroll
                    for (int r = 0; r < 8; ++r)
                        a2[r] = ffma2(make_float2(b[r][2 * j], b[r][2 * j + 1]), make_float2(s[2 * j], s[2 * j + 1]), a2[r]);
            }
        }
    }
    float tot = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) tot += a[r] + a2[r].x + a2[r].y;
    out[blockIdx.x * 128 + threadIdx.x] = tot;
}

template <int MODE, bool PACKED>
void run(const char *name, const float *src, float *out)
{
    const int iters = 4000, blocks = 148 * 3;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, PACKED><<<blocks, 128>>>(src, out, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE, PACKED><<<blocks, 128>>>(src, out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 96 * iters * (double)blocks * 128;
    printf("%-44s %.3f ms  %.1f TFLOP/s  %.1f%% of nominal\n", name, ms, fl / (ms * 1e-3) / 1e12,
           100 * fl / (ms * 1e-3) / (148.0 * 128 * 2 * 1.965e9));
}

int main()
{
    float *src, *out;
    cudaMalloc(&src, 8 * 12 * 128 * 4 + 4096);
    cudaMemset(src, 0, 8 * 12 * 128 * 4 + 4096);
    cudaMalloc(&out, 148 * 3 * 128 * 4);
    run<0, false>("scalar a=fma(a,s,t)             1 fresh", src, out);
    run<1, false>("scalar a[k]=fma(b[k],s,a[k])    2 fresh/8 regs", src, out);
    run<2, false>("scalar a[k]=fma(b[k][j],s[j],a[k]) 2 fresh/96", src, out);
    run<0, true>("packed a=fma(a,s,t)             1 fresh", src, out);
    run<1, true>("packed a[k]=fma(b[k],s,a[k])    2 fresh/8 regs", src, out);
    run<2, true>("packed a[k]=fma(b[k][j],s[j],a[k]) 2 fresh/96", src, out);
    return 0;
}
