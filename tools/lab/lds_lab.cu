// Microbenchmark (development aid): cost of the per-row shared-memory loads next to the FFMA stream.
// Same FP work per row (8 chains x 13 taps + 8 V-FMAs); the 24 operand words arrive as 24 x LDS.32,
// 12 x LDS.64 or 6 x LDS.128 (per-lane contiguous, conflict-free), or are not loaded at all.
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"
namespace tai { void set_error(const char *, ...) {} void count_launch(int) {} }
using namespace tai;

template <int WIDTH, int PACKED>  // WIDTH: 0 none, 1/2/4 words per LDS
__global__ void __launch_bounds__(128, 3) k(const float *hsrc, float *out, int rows)
{
    constexpr int J = 13;
    __shared__ __align__(16) float sm[128 * 24 * 2];
    for (int i = threadIdx.x; i < 128 * 24 * 2; i += 128) sm[i] = 0.001f * i;
    __syncthreads();
    float h[8][J];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int j = 0; j < J; ++j) h[r][j] = hsrc[(r * J + j) * 128 + threadIdx.x];
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float w[24];
#pragma unroll
    for (int j = 0; j < 24; ++j) w[j] = sm[j];
#pragma unroll 1
    for (int it = 0; it < rows; ++it) {
        // lane-contiguous layout: word j of lane l at ((it&1)*24*128) + (j/WIDTH)*128*WIDTH + l*WIDTH + j%WIDTH
        const float *base = sm + (it & 1) * 24 * 128;
        if (WIDTH == 1) {
#pragma unroll
            for (int j = 0; j < 24; ++j) w[j] = base[j * 128 + threadIdx.x];
        } else if (WIDTH == 2) {
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                float2 t = *reinterpret_cast<const float2 *>(base + j * 256 + threadIdx.x * 2);
                w[2 * j] = t.x; w[2 * j + 1] = t.y;
            }
        } else if (WIDTH == 4) {
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                float4 t = *reinterpret_cast<const float4 *>(base + j * 512 + threadIdx.x * 4);
                w[4 * j] = t.x; w[4 * j + 1] = t.y; w[4 * j + 2] = t.z; w[4 * j + 3] = t.w;
            }
        } else {
            w[it & 7] += 1.0f;
        }
        if (PACKED == 3) {
            // row-packed, two independent FFMA2 chains per row pair (even / odd taps) -> 8 chains in flight
#pragma unroll
            for (int rp = 0; rp < 4; ++rp) {
                float2 sa = fmul2(make_float2(h[2 * rp][0], h[2 * rp + 1][0]), make_float2(w[0], w[0]));
                float2 sb = fmul2(make_float2(h[2 * rp][1], h[2 * rp + 1][1]), make_float2(w[1], w[1]));
#pragma unroll
                for (int j = 2; j < J; ++j) {
                    if (j & 1) sb = ffma2(make_float2(h[2 * rp][j], h[2 * rp + 1][j]), make_float2(w[j], w[j]), sb);
                    else sa = ffma2(make_float2(h[2 * rp][j], h[2 * rp + 1][j]), make_float2(w[j], w[j]), sa);
                }
                const float2 vv = make_float2(w[13 + 2 * rp], w[14 + 2 * rp]);
                float2 a2 = ffma2(vv, sa, make_float2(acc[2 * rp], acc[2 * rp + 1]));
                a2 = ffma2(vv, sb, a2);
                acc[2 * rp] = a2.x; acc[2 * rp + 1] = a2.y;
            }
        } else if (PACKED == 2) {
#pragma unroll
            for (int rp = 0; rp < 4; ++rp) {
                float2 s2 = fmul2(make_float2(h[2 * rp][0], h[2 * rp + 1][0]), make_float2(w[0], w[0]));
#pragma unroll
                for (int j = 1; j < J; ++j)
                    s2 = ffma2(make_float2(h[2 * rp][j], h[2 * rp + 1][j]), make_float2(w[j], w[j]), s2);
                float2 a2 = ffma2(make_float2(w[13 + 2 * rp], w[14 + 2 * rp]), s2, make_float2(acc[2 * rp], acc[2 * rp + 1]));
                acc[2 * rp] = a2.x; acc[2 * rp + 1] = a2.y;
            }
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float s = h[r][0] * w[0];
#pragma unroll
                for (int j = 1; j < J; ++j) s = fmaf(h[r][j], w[j], s);
                acc[r] = fmaf(w[13 + r], s, acc[r]);
            }
        }
    }
    float t = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += acc[r];
    out[blockIdx.x * 128 + threadIdx.x] = t;
}

template <int WIDTH, int PACKED>
void run(const char *name, const float *h, float *out)
{
    const int rows = 4000, blocks = 148 * 3;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<WIDTH, PACKED><<<blocks, 128>>>(h, out, rows);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<WIDTH, PACKED><<<blocks, 128>>>(h, out, rows);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 8 * 13 * rows * (double)blocks * 128;
    printf("%-34s %.3f ms  %.1f%% of nominal  [%.0f cycles per SMSP row-round of 3 warps]\n", name, ms,
           100 * fl / (ms * 1e-3) / (148.0 * 128 * 2 * 1.965e9), ms * 1e-3 * 1.92e9 / rows);
}

int main()
{
    float *h, *out;
    cudaMalloc(&h, 8 * 16 * 128 * 4);
    cudaMemset(h, 0, 8 * 16 * 128 * 4);
    cudaMalloc(&out, 148 * 4 * 128 * 4);
    run<0, 0>("scalar, no LDS", h, out);
    run<1, 0>("scalar, 24 x LDS.32", h, out);
    run<2, 0>("scalar, 12 x LDS.64", h, out);
    run<4, 0>("scalar,  6 x LDS.128", h, out);
    run<1, 3>("rowpacked 8 chains, 24 x LDS.32", h, out);
    run<4, 3>("rowpacked 8 chains, 6 x LDS.128", h, out);
    run<0, 2>("rowpacked, no LDS", h, out);
    run<1, 2>("rowpacked, 24 x LDS.32", h, out);
    run<2, 2>("rowpacked, 12 x LDS.64", h, out);
    run<4, 2>("rowpacked,  6 x LDS.128", h, out);
    return 0;
}
