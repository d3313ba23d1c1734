// Microbenchmark (development aid): why do the kernels that ACCUMULATE INTO 104 REGISTERS (gH: a[r][k] += w[r]*iv[k])
// issue slower than the forward sweep (s[r] += h[r][k]*iv[k], 104 read-only taps, 8 accumulators)?  Same FFMA count,
// same shared-memory reads, 3 CTAs x 4 warps per SM, no global traffic.
//   VAR 0: forward body (tap-outer)                          104 FFMA + 8 FFMA + 21 LDS per row
//   VAR 1: gH body, tap-outer / row-inner                    104 FFMA + 8 FMUL + 21 LDS per row
//   VAR 2: gH body, row-outer / tap-inner
//   VAR 3: forward body, registers only (no LDS)             multiplicands from a rotating register set
//   VAR 4: gH body, registers only
//   VAR 5: gH body with packed FFMA2 over row pairs: (a[r][k], a[r+1][k]) += (w[r], w[r+1]) * (iv[k], iv[k])
//   VAR 6: gH body with packed FFMA2 over tap pairs: (a[r][k], a[r][k+1]) += (w[r], w[r]) * (iv[k], iv[k+1])
//   nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -I video_frame_inpainting_b200/csrc \
//        tools/lab/acc_lab.cu -o tools/lab/acc_lab
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"
namespace tai { void set_error(const char *, ...) {} void count_launch(int) {} void note_path(const char *) {} }
using namespace tai;

constexpr int KS = 51, P = 8, J = 13, TW = 32, PITCH = TW + 52, ROWS = P + KS - 1, VROW = P * TW;

__device__ __forceinline__ unsigned long long pack2(float x, float y)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ void ffma2_acc(unsigned long long &acc, unsigned long long a, unsigned long long b)
{
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ float sum2(unsigned long long v)
{
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
    return x + y;
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
    const float *vrow0 = slab + warp * 8 + cx;
    const float *srow0 = is + warp * 8 + cx + ch;
    float total = 0.f;

    if (VAR == 0 || VAR == 3) {
        float h[P][J], acc[P];
#pragma unroll
        for (int r = 0; r < P; ++r) {
            acc[r] = 0.f;
#pragma unroll
            for (int j = 0; j < J; ++j) h[r][j] = hsrc[(r * 16 + j) * 128 + threadIdx.x];
        }
        float iv[J], v[P];
#pragma unroll
        for (int j = 0; j < J; ++j) iv[j] = hsrc[j * 128 + threadIdx.x] + 1.f;
#pragma unroll
        for (int r = 0; r < P; ++r) v[r] = hsrc[(r + 3) * 128 + threadIdx.x] + 2.f;
#pragma unroll 1
        for (int it = 0; it < sweeps; ++it) {
#pragma unroll 1
            for (int yy = P - 1; yy < KS; ++yy) {
                if (VAR == 0) {
                    const float *srow = srow0 + yy * PITCH;
                    const float *vrow = vrow0 + yy * VROW;
#pragma unroll
                    for (int r = 0; r < P; ++r) v[r] = vrow[r * (TW - VROW)];
#pragma unroll
                    for (int j = 0; j < J; ++j) iv[j] = srow[4 * j];
                } else {   // rotate the multiplicands so that nothing is loop invariant
                    const float t0 = iv[0];
#pragma unroll
                    for (int j = 0; j + 1 < J; ++j) iv[j] = iv[j + 1];
                    iv[J - 1] = t0;
                }
                float s[P];
#pragma unroll
                for (int r = 0; r < P; ++r) s[r] = h[r][0] * iv[0];
#pragma unroll
                for (int j = 1; j < J; ++j)
#pragma unroll
                    for (int r = 0; r < P; ++r) s[r] = fmaf(h[r][j], iv[j], s[r]);
#pragma unroll
                for (int r = 0; r < P; ++r) acc[r] = fmaf(v[r], s[r], acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < P; ++r) total += acc[r];
    } else if (VAR == 1 || VAR == 2 || VAR == 4) {
        float a[P][J], go[P];
#pragma unroll
        for (int r = 0; r < P; ++r) {
            go[r] = hsrc[r * 128 + threadIdx.x] + 1.f;
#pragma unroll
            for (int j = 0; j < J; ++j) a[r][j] = 0.f;
        }
        float iv[J], v[P];
#pragma unroll
        for (int j = 0; j < J; ++j) iv[j] = hsrc[j * 128 + threadIdx.x] + 1.f;
#pragma unroll
        for (int r = 0; r < P; ++r) v[r] = hsrc[(r + 3) * 128 + threadIdx.x] + 2.f;
#pragma unroll 1
        for (int it = 0; it < sweeps; ++it) {
#pragma unroll 1
            for (int yy = P - 1; yy < KS; ++yy) {
                if (VAR != 4) {
                    const float *srow = srow0 + yy * PITCH;
                    const float *vrow = vrow0 + yy * VROW;
#pragma unroll
                    for (int r = 0; r < P; ++r) v[r] = vrow[r * (TW - VROW)];
#pragma unroll
                    for (int j = 0; j < J; ++j) iv[j] = srow[4 * j];
                } else {
                    const float t0 = iv[0];
#pragma unroll
                    for (int j = 0; j + 1 < J; ++j) iv[j] = iv[j + 1];
                    iv[J - 1] = t0;
                }
                float w[P];
#pragma unroll
                for (int r = 0; r < P; ++r) w[r] = v[r] * go[r];
                if (VAR == 2) {
#pragma unroll
                    for (int r = 0; r < P; ++r)
#pragma unroll
                        for (int j = 0; j < J; ++j) a[r][j] = fmaf(w[r], iv[j], a[r][j]);
                } else {
#pragma unroll
                    for (int j = 0; j < J; ++j)
#pragma unroll
                        for (int r = 0; r < P; ++r) a[r][j] = fmaf(w[r], iv[j], a[r][j]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < P; ++r)
#pragma unroll
            for (int j = 0; j < J; ++j) total += a[r][j];
    } else if (VAR == 5) {   // FFMA2 over row pairs
        unsigned long long a2[P / 2][J];
        float go[P];
#pragma unroll
        for (int r = 0; r < P; ++r) go[r] = hsrc[r * 128 + threadIdx.x] + 1.f;
#pragma unroll
        for (int q = 0; q < P / 2; ++q)
#pragma unroll
            for (int j = 0; j < J; ++j) a2[q][j] = 0ull;
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
                unsigned long long w2[P / 2];
#pragma unroll
                for (int q = 0; q < P / 2; ++q) w2[q] = pack2(v[2 * q] * go[2 * q], v[2 * q + 1] * go[2 * q + 1]);
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const unsigned long long d = pack2(iv[j], iv[j]);
#pragma unroll
                    for (int q = 0; q < P / 2; ++q) ffma2_acc(a2[q][j], w2[q], d);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < P / 2; ++q)
#pragma unroll
            for (int j = 0; j < J; ++j) total += sum2(a2[q][j]);
    } else {                 // VAR 6: FFMA2 over tap pairs (7 pairs: 14 slots, the last one padding)
        constexpr int JP = 7;
        unsigned long long a2[P][JP];
        float go[P];
#pragma unroll
        for (int r = 0; r < P; ++r) go[r] = hsrc[r * 128 + threadIdx.x] + 1.f;
#pragma unroll
        for (int r = 0; r < P; ++r)
#pragma unroll
            for (int j = 0; j < JP; ++j) a2[r][j] = 0ull;
#pragma unroll 1
        for (int it = 0; it < sweeps; ++it) {
#pragma unroll 1
            for (int yy = P - 1; yy < KS; ++yy) {
                const float *srow = srow0 + yy * PITCH;
                const float *vrow = vrow0 + yy * VROW;
                float v[P];
#pragma unroll
                for (int r = 0; r < P; ++r) v[r] = vrow[r * (TW - VROW)];
                unsigned long long iv2[JP];
#pragma unroll
                for (int j = 0; j < JP; ++j) iv2[j] = pack2(srow[8 * j], srow[8 * j + 4]);
                unsigned long long w2[P];
#pragma unroll
                for (int r = 0; r < P; ++r) {
                    const float w = v[r] * go[r];
                    w2[r] = pack2(w, w);
                }
#pragma unroll
                for (int j = 0; j < JP; ++j)
#pragma unroll
                    for (int r = 0; r < P; ++r) ffma2_acc(a2[r][j], w2[r], iv2[j]);
            }
        }
#pragma unroll
        for (int r = 0; r < P; ++r)
#pragma unroll
            for (int j = 0; j < JP; ++j) total += sum2(a2[r][j]);
    }
    out[blockIdx.x * 128 + threadIdx.x] = total;
}

template <int VAR>
void run(const char *name, const float *h, float *out)
{
    const int sweeps = 100, blocks = 148 * 3;
    const size_t smem = (KS * VROW + ROWS * PITCH) * 4;
    cudaFuncSetAttribute(k<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<VAR>, 128, smem);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k<VAR>);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<VAR><<<blocks, 128, smem>>>(h, out, sweeps);
    cudaDeviceSynchronize();
    float best = 1e9;
    for (int i = 0; i < 3; ++i) {
        cudaEventRecord(e0);
        k<VAR><<<blocks, 128, smem>>>(h, out, sweeps);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const int rows = sweeps * (KS - P + 1);
    const double fl = 2.0 * 104 * rows * (double)blocks * 128;   // 104 FMAs per row and lane in every variant
    printf("%-52s regs=%3d occ=%d %.3f ms  %.1f%% of nominal FMA peak  [%.0f cycles per warp-row at 3 warps/SMSP]\n", name,
           fa.numRegs, occ, best, 100 * fl / (best * 1e-3) / (148.0 * 128 * 2 * 1.965e9), best * 1e-3 * 1.93e9 / rows / 3);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
}

int main()
{
    float *h, *out;
    cudaMalloc(&h, 8 * 16 * 128 * 4);
    cudaMemset(h, 0, 8 * 16 * 128 * 4);
    cudaMalloc(&out, 148 * 4 * 128 * 4);
    run<0>("forward body: s[r] += h[r][k]*iv[k]", h, out);
    run<1>("gH body: a[r][k] += w[r]*iv[k], tap-outer", h, out);
    run<2>("gH body, row-outer", h, out);
    run<3>("forward body, registers only", h, out);
    run<4>("gH body, registers only", h, out);
    run<5>("gH body, FFMA2 over row pairs", h, out);
    run<6>("gH body, FFMA2 over tap pairs", h, out);
    return 0;
}
