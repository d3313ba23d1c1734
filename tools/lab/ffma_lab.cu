// Microbenchmark (development aid): throughput of the sweep's inner product with register-resident H
// taps, scalar FFMA vs packed FFMA2, with and without the per-row shared-memory loads.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "common.cuh"
namespace tai { void set_error(const char *, ...) {} void count_launch(int) {} }
using namespace tai;

template <int J, int PACKED, bool WITH_LDS>
__global__ void __launch_bounds__(128, 3) k(const float *hsrc, float *out, int rows)
{
    __shared__ float sm[4096 + 512];
    for (int i = threadIdx.x; i < 4096 + 512; i += 128) sm[i] = 0.001f * i;
    __syncthreads();
    float h[8][J];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int j = 0; j < J; ++j) h[r][j] = hsrc[(r * J + j) * 128 + threadIdx.x];
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int lo = (threadIdx.x & 7) + ((threadIdx.x >> 3) & 3);
    float iv[J], v[8];
#pragma unroll
    for (int j = 0; j < J; ++j) iv[j] = sm[lo + 4 * j];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = sm[lo + r * 32];
#pragma unroll 1
    for (int it = 0; it < rows; ++it) {
        if (WITH_LDS) {
            const float *row = sm + ((it * 84) & 4095) + lo;
#pragma unroll
            for (int j = 0; j < J; ++j) iv[j] = row[4 * j];
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r] = row[r * 32 + 3];
        } else {
            iv[it & 1] += 1.0f;  // keep the products loop-variant
        }
        if (PACKED == 2) {
            // pack over output rows: (s[r], s[r+1]) += (h[r][j], h[r+1][j]) * (iv[j], iv[j])
#pragma unroll
            for (int rp = 0; rp < 4; ++rp) {
                float2 s2 = fmul2(make_float2(h[2 * rp][0], h[2 * rp + 1][0]), make_float2(iv[0], iv[0]));
#pragma unroll
                for (int j = 1; j < J; ++j)
                    s2 = ffma2(make_float2(h[2 * rp][j], h[2 * rp + 1][j]), make_float2(iv[j], iv[j]), s2);
                float2 a2 = ffma2(make_float2(v[2 * rp], v[2 * rp + 1]), s2, make_float2(acc[2 * rp], acc[2 * rp + 1]));
                acc[2 * rp] = a2.x;
                acc[2 * rp + 1] = a2.y;
            }
        } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float s;
            if (PACKED) {
                float2 s2 = fmul2(make_float2(h[r][0], h[r][1]), make_float2(iv[0], iv[1]));
#pragma unroll
                for (int q = 1; q < J / 2; ++q)
                    s2 = ffma2(make_float2(h[r][2 * q], h[r][2 * q + 1]), make_float2(iv[2 * q], iv[2 * q + 1]), s2);
                s = s2.x + s2.y;
                if (J & 1) s = fmaf(h[r][J - 1], iv[J - 1], s);
            } else {
                s = h[r][0] * iv[0];
#pragma unroll
                for (int j = 1; j < J; ++j) s = fmaf(h[r][j], iv[j], s);
            }
            acc[r] = fmaf(v[r], s, acc[r]);
        }
        }
    }
    float t = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += acc[r];
    out[blockIdx.x * 128 + threadIdx.x] = t;
}

template <int J, int PACKED, bool WITH_LDS>
void run(const char *name, const float *h, float *out, int blocks)
{
    const int rows = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<J, PACKED, WITH_LDS><<<blocks, 128>>>(h, out, rows);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<J, PACKED, WITH_LDS><<<blocks, 128>>>(h, out, rows);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 8 * J * rows * (double)blocks * 128;
    double cyc_per_row = ms * 1e-3 * 1.92e9 / rows / (blocks / 148.0 / 1.0) * 1.0;  // per SM-row-batch
    printf("%-28s blocks/SM=%d  %.3f ms  %.1f TFLOP/s (taps only)  %.1f%% of nominal  [%.0f cyc per row per block-set]\n", name,
           blocks / 148, ms, fl / (ms * 1e-3) / 1e12, 100 * fl / (ms * 1e-3) / (148.0 * 128 * 2 * 1.965e9), cyc_per_row);
}

int main()
{
    float *h, *out;
    cudaMalloc(&h, 8 * 16 * 128 * 4);
    cudaMemset(h, 0, 8 * 16 * 128 * 4);
    cudaMalloc(&out, 148 * 4 * 128 * 4);
    for (int bps = 2; bps <= 3; ++bps) {
        run<13, 0, true>("scalar J=13 + 21 LDS/row", h, out, 148 * bps);
        run<13, 1, true>("packed(j) J=13 + 21 LDS/row", h, out, 148 * bps);
        run<13, 2, true>("packed(rows) J=13 + 21 LDS/row", h, out, 148 * bps);
    }
    return 0;
}
