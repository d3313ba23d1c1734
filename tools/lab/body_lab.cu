// Microbenchmark (development aid): the row body of the forward sweep with the REAL shared-memory
// addressing (halo [ROWS][PITCH], V slab [tap][8][32]) but no global traffic, 3 CTAs x 4 warps per SM.
//   VAR 0: scalar FFMA, taps j == ch (mod 4)                       (the v3 body)
//   VAR 1: FFMA2 packed over adjacent taps: lane group ch owns the pairs {8k+2ch, 8k+2ch+1}, k < 6, and the
//          single tap 48+ch; the halo pair arrives as two LDS.32 into an aligned register pair
//   VAR 2: as 1, all 7 slots packed (taps 8k+2ch+{0,1}, k < 7; taps >= 51 are zero)
//   nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -I video_frame_inpainting_b200/csrc \
//        tools/lab/body_lab.cu -o tools/lab/body_lab
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"
namespace tai { void set_error(const char *, ...) {} void count_launch(int) {} }
using namespace tai;

constexpr int KS = 51, P = 8, TW = 32, PITCH = TW + 52, ROWS = P + KS - 1, VROW = P * TW;

__device__ __forceinline__ float2 ffma2r(float2 a, float2 b, float2 c)
{
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
__device__ __forceinline__ float2 fmul2r(float2 a, float2 b)
{
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}

template <int VAR>
__global__ void __launch_bounds__(128, 3) k(const float *hsrc, float *out, int sweeps)
{
    extern __shared__ __align__(16) float sm[];
    float *slab = sm;                 // [KS][P][TW]
    float *is = sm + KS * VROW;       // [ROWS][PITCH]
    for (int i = threadIdx.x; i < KS * VROW + ROWS * PITCH; i += 128) sm[i] = 1e-3f * (i % 977);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, cx = lane & 7, ch = lane >> 3;
    float acc[P];
#pragma unroll
    for (int r = 0; r < P; ++r) acc[r] = 0.f;
    const float *vrow0 = slab + warp * 8 + cx;

    if (VAR == 0 || VAR == 3 || VAR == 4) {
        constexpr int J = 13;
        float h[P][J];
#pragma unroll
        for (int r = 0; r < P; ++r)
#pragma unroll
            for (int j = 0; j < J; ++j) h[r][j] = hsrc[(r * 16 + j) * 128 + threadIdx.x];
        const float *srow0 = is + warp * 8 + cx + ch;
#pragma unroll 1
        for (int it = 0; it < sweeps; ++it) {
#pragma unroll 1
            for (int yy = P - 1; yy < KS; ++yy) {
                const float *srow = srow0 + yy * PITCH;
                const float *vrow = vrow0 + yy * VROW;
                float v[P], iv[J];
#pragma unroll
                for (int r = 0; r < P; ++r) v[r] = vrow[r * (TW - VROW)];
#pragma unroll
                for (int j = 0; j < J; ++j) iv[j] = srow[4 * j];
                if (VAR == 0) {
#pragma unroll
                    for (int r = 0; r < P; ++r) {
                        float s = h[r][0] * iv[0];
#pragma unroll
                        for (int j = 1; j < J; ++j) s = fmaf(h[r][j], iv[j], s);
                        acc[r] = fmaf(v[r], s, acc[r]);
                    }
                } else if (VAR == 3) {   // tap-outer order: 8 independent chains interleaved
                    float s[P];
#pragma unroll
                    for (int r = 0; r < P; ++r) s[r] = h[r][0] * iv[0];
#pragma unroll
                    for (int j = 1; j < J; ++j)
#pragma unroll
                        for (int r = 0; r < P; ++r) s[r] = fmaf(h[r][j], iv[j], s[r]);
#pragma unroll
                    for (int r = 0; r < P; ++r) acc[r] = fmaf(v[r], s[r], acc[r]);
                } else {                 // tap-outer, V folded in first: s[r] = v[r]*h0*iv0 ... no extra FMA: acc += v*(sum)
                    float s[P];
#pragma unroll
                    for (int r = 0; r < P; ++r) s[r] = h[r][0] * iv[0];
#pragma unroll
                    for (int j = 1; j < J; ++j) {
#pragma unroll
                        for (int r = 0; r < P; ++r) s[r] = fmaf(h[r][j], iv[j], s[r]);
                        asm volatile("" ::: "memory");
                    }
#pragma unroll
                    for (int r = 0; r < P; ++r) acc[r] = fmaf(v[r], s[r], acc[r]);
                }
            }
        }
    } else {
        constexpr int NP = (VAR == 1) ? 6 : 7;
        float2 h2[P][NP];
        float h1[P];
#pragma unroll
        for (int r = 0; r < P; ++r) {
#pragma unroll
            for (int j = 0; j < NP; ++j)
                h2[r][j] = make_float2(hsrc[(r * 16 + 2 * j) * 128 + threadIdx.x], hsrc[(r * 16 + 2 * j + 1) * 128 + threadIdx.x]);
            h1[r] = hsrc[(r * 16 + 14) * 128 + threadIdx.x];
        }
        const float *srow0 = is + warp * 8 + cx + 2 * ch;
        const float *srow1 = is + warp * 8 + cx + 48 + ch;
#pragma unroll 1
        for (int it = 0; it < sweeps; ++it) {
#pragma unroll 1
            for (int yy = P - 1; yy < KS; ++yy) {
                const float *srow = srow0 + yy * PITCH;
                const float *vrow = vrow0 + yy * VROW;
                float v[P];
                float2 iv2[NP];
                float iv1 = 0.f;
#pragma unroll
                for (int r = 0; r < P; ++r) v[r] = vrow[r * (TW - VROW)];
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    iv2[j].x = srow[8 * j];
                    iv2[j].y = srow[8 * j + 1];
                }
                if (VAR == 1) iv1 = srow1[yy * PITCH];
#pragma unroll
                for (int r = 0; r < P; ++r) {
                    float2 s2 = fmul2r(h2[r][0], iv2[0]);
#pragma unroll
                    for (int j = 1; j < NP; ++j) s2 = ffma2r(h2[r][j], iv2[j], s2);
                    float s = s2.x + s2.y;
                    if (VAR == 1) s = fmaf(h1[r], iv1, s);
                    acc[r] = fmaf(v[r], s, acc[r]);
                }
            }
        }
    }
    float t = 0;
#pragma unroll
    for (int r = 0; r < P; ++r) t += acc[r];
    out[blockIdx.x * 128 + threadIdx.x] = t;
}

template <int VAR>
void run(const char *name, const float *h, float *out)
{
    const int sweeps = 100, blocks = 148 * 3;
    const size_t smem = (KS * VROW + ROWS * PITCH) * 4;
    cudaFuncSetAttribute(k<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<VAR>, 128, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<VAR><<<blocks, 128, smem>>>(h, out, sweeps);
    cudaDeviceSynchronize();
    float best = 1e9;
    for (int i = 0; i < 3; ++i) {
        cudaEventRecord(e0);
        k<VAR><<<blocks, 128, smem>>>(h, out, sweeps);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const int rows = sweeps * (KS - P + 1);
    double fl = 2.0 * 8 * 51 / 4.0 * rows * (double)blocks * 128;   // useful flop: 51 taps over 4 lane groups
    printf("%-40s occ=%d %.3f ms  %.1f%% of nominal (useful taps)  [%.0f cycles per warp-row at 3 warps/SMSP]\n", name, occ,
           best, 100 * fl / (best * 1e-3) / (148.0 * 128 * 2 * 1.965e9), best * 1e-3 * 1.93e9 / rows / 3);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
}

int main()
{
    float *h, *out;
    cudaMalloc(&h, 8 * 16 * 128 * 4);
    cudaMemset(h, 0, 8 * 16 * 128 * 4);
    cudaMalloc(&out, 148 * 4 * 128 * 4);
    run<0>("scalar (v3 body)", h, out);
    run<1>("FFMA2 tap pairs 6+1", h, out);
    run<2>("FFMA2 tap pairs 7", h, out);
    run<3>("scalar, tap-outer order", h, out);
    run<4>("scalar, tap-outer order, fenced", h, out);
    return 0;
}
