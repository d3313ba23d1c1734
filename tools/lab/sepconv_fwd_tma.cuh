// Forward separable convolution, TMA-staged variant (compile-time ks).
//
// Same lane layout as sepconv_fwd_kernel (sepconv_fwd.cu): a warp owns 8 columns x 8 rows, lane group
// ch = lane>>3 owns horizontal taps j == ch (mod 4), H taps live in registers, I comes from the
// shared-memory halo.  What changes:
//   * the block's slab of the VERTICAL kernel map, V[b, 0..ks, y0..y0+8, x0..x0+32] (ks KB), is brought
//     into shared memory by ONE cp.async.bulk.tensor (TMA) issued by one thread and awaited on an
//     mbarrier; it overlaps the halo staging and the H register loads.  In the sweep V is then an LDS
//     with an immediate offset instead of an HBM-latency-exposed LDG with 64-bit address arithmetic;
//   * ks is a template parameter, so every shared-memory offset in the sweep is an immediate;
//   * PACKED: the inner products run on FFMA2 (fma.rn.f32x2): two taps per issue slot.
#pragma once

#include "common.cuh"
#include "sepconv_common.cuh"
#include "tma.cuh"

namespace tai {

#ifdef TAI_LAB_TIMING
__device__ unsigned long long g_lab_phase[8];
#define LAB_T(i)                                                  \
    do {                                                          \
        if (threadIdx.x == 0) {                                   \
            const long long now_ = clock64();                     \
            atomicAdd(&g_lab_phase[i], (unsigned long long)(now_ - lab_t_)); \
            lab_t_ = now_;                                        \
        }                                                         \
    } while (0)
#else
#define LAB_T(i)
#endif

template <int KS>
struct FwdTmaCfg {
    static constexpr int J = (KS + 3) / 4;
    static constexpr int WX = 4;
    static constexpr int NT = 32 * WX;
    static constexpr int TILE_W = WX * FNX, TILE_H = FP;
    static constexpr int PITCH = TILE_W + 4 * J;
    static constexpr int ROWS = TILE_H + KS - 1;
    static constexpr int VS_FLOATS = KS * TILE_H * TILE_W;
    static constexpr int VROW = TILE_H * TILE_W;  // floats between consecutive taps in the V slab
    static constexpr size_t smem_bytes(int cg) { return (size_t)(VS_FLOATS + cg * ROWS * PITCH) * 4 + 16; }
};

// Operands of one input row of the sweep: the lane's J halo words per channel and its P vertical taps.
template <int KS, int CG>
struct RowRegs {
    float iv[CG][(KS + 3) / 4];
    float v[FP];
};

template <int KS, int CG, int RLO, int RHI>
__device__ __forceinline__ void fwd_row_load(const float *__restrict__ srow, const float *__restrict__ vrow,
                                             RowRegs<KS, CG> &rr)
{
    using Cfg = FwdTmaCfg<KS>;
    constexpr int CSTRIDE = Cfg::ROWS * Cfg::PITCH;
#pragma unroll
    for (int r = RLO; r < RHI; ++r) rr.v[r] = vrow[r * (Cfg::TILE_W - Cfg::VROW)];  // tap yy-r, row r
#pragma unroll
    for (int c = 0; c < CG; ++c)
#pragma unroll
        for (int jj = 0; jj < Cfg::J; ++jj) rr.iv[c][jj] = srow[c * CSTRIDE + 4 * jj];
}

template <int KS, int CG, int RLO, int RHI, bool PACKED>
__device__ __forceinline__ void fwd_row_math(const RowRegs<KS, CG> &rr, const float (&h)[FP][(KS + 3) / 4],
                                             float (&acc)[CG][FP])
{
    constexpr int J = (KS + 3) / 4;
#pragma unroll
    for (int c = 0; c < CG; ++c) {
#pragma unroll
        for (int r = RLO; r < RHI; ++r) {
            float s;
            if (PACKED) {
                float2 s2 = fmul2(make_float2(h[r][0], h[r][1]), make_float2(rr.iv[c][0], rr.iv[c][1]));
#pragma unroll
                for (int q = 1; q < J / 2; ++q)
                    s2 = ffma2(make_float2(h[r][2 * q], h[r][2 * q + 1]),
                               make_float2(rr.iv[c][2 * q], rr.iv[c][2 * q + 1]), s2);
                s = s2.x + s2.y;
                if (J & 1) s = fmaf(h[r][J - 1], rr.iv[c][J - 1], s);
            } else {
                s = h[r][0] * rr.iv[c][0];
#pragma unroll
                for (int jj = 1; jj < J; ++jj) s = fmaf(h[r][jj], rr.iv[c][jj], s);
            }
            acc[c][r] = fmaf(rr.v[r], s, acc[c][r]);
        }
    }
}

template <int KS, int CG, int RLO, int RHI, bool PACKED>
__device__ __forceinline__ void fwd_row_tma(const float *__restrict__ srow, const float *__restrict__ vrow,
                                            const float (&h)[FP][(KS + 3) / 4], float (&acc)[CG][FP])
{
    RowRegs<KS, CG> rr;
    fwd_row_load<KS, CG, RLO, RHI>(srow, vrow, rr);
    fwd_row_math<KS, CG, RLO, RHI, PACKED>(rr, h, acc);
}

template <int KS, int CG, bool PAD, bool DUAL, bool PACKED, bool PIPE = false, int MINB = (CG == 1 ? 3 : 2)>
__global__ void __launch_bounds__(128, MINB)
sepconv_fwd_tma_kernel(const __grid_constant__ CUtensorMap tmv0, const __grid_constant__ CUtensorMap tmv1,
                       const __grid_constant__ CUtensorMap tmh0, const __grid_constant__ CUtensorMap tmh1,
                       const FwdParams p)
{
    using Cfg = FwdTmaCfg<KS>;
    constexpr int J = Cfg::J, PITCH = Cfg::PITCH, ROWS = Cfg::ROWS, TILE_W = Cfg::TILE_W, TILE_H = Cfg::TILE_H;
    constexpr int CSTRIDE = ROWS * PITCH;
    extern __shared__ __align__(128) float smem[];
    float *vs = smem;                       // [KS][TILE_H][TILE_W], written by TMA
    float *is = smem + Cfg::VS_FLOATS;      // [CG][ROWS][PITCH]
    uint64_t *bar = reinterpret_cast<uint64_t *>(is + CG * CSTRIDE);

    const int Ho = p.Ho, Wo = p.Wo;
    const int Hi = Ho + KS - 1, Wi = Wo + KS - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cx = lane & 7, ch = lane >> 3;

    const long plane = (long)Ho * Wo;
    const int ntiles = p.B * p.nty * p.ntx;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    uint32_t phase = 0;

    // Persistent CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ...  While tile k is being filtered the
    // kernel-map slabs of tile k+1 are pulled into L2 by two TMA prefetches, so that the next
    // iteration's TMA / LDG traffic is served by L2 and the HBM stream overlaps the FMA work.
#ifdef TAI_LAB_TIMING
    long long lab_t_ = clock64();
#endif
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int t = tile;
    const int tx = t % p.ntx;
    t /= p.ntx;
    const int ty = t % p.nty;
    const int b = t / p.nty;
    const int x0 = max(0, min(tx * TILE_W, Wo - TILE_W));
    const int y0 = min(ty * TILE_H, Ho - TILE_H);  // host guarantees Ho >= TILE_H
    const int px_raw = x0 + warp * FNX + cx;
    const bool px_ok = px_raw < Wo;
    const int px = px_ok ? px_raw : Wo - 1;
    if (threadIdx.x == 32 && tile + (int)gridDim.x < ntiles) {
        int n = tile + gridDim.x;
        const int ntx_ = n % p.ntx;
        n /= p.ntx;
        const int nty_ = n % p.nty;
        const int nb = n / p.nty;
        const int nx0 = max(0, min(ntx_ * TILE_W, Wo - TILE_W)), ny0 = min(nty_ * TILE_H, Ho - TILE_H);
        tma_prefetch_l2_4d(&tmv0, nx0, ny0, 0, nb);
        tma_prefetch_l2_4d(&tmh0, nx0, ny0, 0, nb);
        if (DUAL) {
            tma_prefetch_l2_4d(&tmv1, nx0, ny0, 0, nb);
            tma_prefetch_l2_4d(&tmh1, nx0, ny0, 0, nb);
        }
    }
    float res[DUAL ? 2 : 1][CG][FP];

    for (int c0 = 0; c0 < p.C; c0 += CG) {
#pragma unroll
        for (int s = 0; s < (DUAL ? 2 : 1); ++s) {
            const float *__restrict__ in = p.in[s];
            const float *__restrict__ hor = p.hor[s];

            // ---- kick off the V slab (TMA), then stage the halo and the H taps meanwhile ----
#ifndef TAI_LAB_SKIP_LOADS
            if (threadIdx.x == 0) {
                fence_proxy_async();  // earlier generic reads of vs are ordered before the refill
                mbar_expect_tx(bar, Cfg::VS_FLOATS * 4);
                tma_load_4d(vs, s == 0 ? &tmv0 : &tmv1, bar, x0, y0, 0, b);
            }
            // halo: one 4-byte cp.async per element (clamped / bounds-checked source), all in flight at once
            for (int c = 0; c < CG; ++c) {
                const float *src = PAD ? in + ((long)(b * p.C + c0 + c)) * plane
                                       : in + ((long)(b * p.C + c0 + c)) * Hi * Wi;
                for (int ry = warp; ry < ROWS; ry += Cfg::NT / 32) {
                    const int gy = y0 + ry;
#pragma unroll
                    for (int k = 0; k < (PITCH + 31) / 32; ++k) {
                        const int rx = lane + 32 * k;
                        if (rx < PITCH) {
                            const int gx = x0 + rx;
                            const float *g;
                            bool valid = rx < TILE_W + KS - 1;
                            if (PAD) {
                                const int sy = clampi(gy - KS / 2, 0, Ho - 1);
                                const int sx = clampi(gx - KS / 2, 0, Wo - 1);
                                g = src + (long)sy * Wo + sx;
                            } else {
                                valid = valid && gx < Wi;
                                g = src + (long)gy * Wi + (valid ? gx : 0);
                            }
                            cp_async_f32(is + c * CSTRIDE + ry * PITCH + rx, g, valid);
                        }
                    }
                }
            }
            cp_async_commit();
#endif
            float h[FP][J];
#ifdef TAI_LAB_SKIP_LOADS
#pragma unroll
            for (int jj = 0; jj < J; ++jj)
#pragma unroll
                for (int r = 0; r < FP; ++r) h[r][jj] = 0.001f * (float)(jj + r + threadIdx.x);
#else
            {
                const float *hp = hor + ((long)b * KS * Ho + y0) * Wo + px;
#pragma unroll
                for (int jj = 0; jj < J; ++jj) {
                    const int j = ch + 4 * jj;
#pragma unroll
                    for (int r = 0; r < FP; ++r)
                        h[r][jj] = (j < KS) ? ld_stream(hp + (long)j * plane + (long)r * Wo) : 0.f;
                }
            }
#endif
#ifndef TAI_LAB_SKIP_LOADS
            LAB_T(0);  // issue of TMA + halo cp.async + H LDGs
            cp_async_wait_all();
            __syncthreads();
            LAB_T(1);  // halo landed
            mbar_wait(bar, phase);
            phase ^= 1;
            LAB_T(2);  // V slab landed
#endif
#ifdef TAI_LAB_SKIP_MATH
            if (p.ks > 0) {
                float a = 0.f;
#pragma unroll
                for (int jj = 0; jj < J; ++jj)
#pragma unroll
                    for (int r = 0; r < FP; ++r) a += h[r][jj];
                a += vs[threadIdx.x * 97 % Cfg::VS_FLOATS] + is[threadIdx.x * 31 % CSTRIDE];
                if (px_ok) p.out[0][((long)(b * p.C) * Ho + y0) * Wo + px] = a;
                return;
            }
#endif

            float acc[CG][FP];
#pragma unroll
            for (int c = 0; c < CG; ++c)
#pragma unroll
                for (int r = 0; r < FP; ++r) acc[c][r] = 0.f;

            const float *srow = is + warp * FNX + cx + ch;
            const float *vrow = vs + warp * FNX + cx;  // tap 0, row 0

            static_for<0, FP - 1>([&](auto YY) {
                constexpr int yy = decltype(YY)::value;
                fwd_row_tma<KS, CG, 0, yy + 1, PACKED>(srow + yy * PITCH, vrow + yy * Cfg::VROW, h, acc);
            });
            if (PIPE && ((KS - (FP - 1)) % 2 == 0)) {
                // software pipeline: the LDS of row yy+1 are in flight while row yy is on the FMA pipe
                RowRegs<KS, CG> ra, rb;
                fwd_row_load<KS, CG, 0, FP>(srow + (FP - 1) * PITCH, vrow + (FP - 1) * Cfg::VROW, ra);
#pragma unroll 1
                for (int yy = FP - 1; yy < KS; yy += 2) {
                    fwd_row_load<KS, CG, 0, FP>(srow + (yy + 1) * PITCH, vrow + (yy + 1) * Cfg::VROW, rb);
                    fwd_row_math<KS, CG, 0, FP, PACKED>(ra, h, acc);
                    if (yy + 2 < KS)
                        fwd_row_load<KS, CG, 0, FP>(srow + (yy + 2) * PITCH, vrow + (yy + 2) * Cfg::VROW, ra);
                    fwd_row_math<KS, CG, 0, FP, PACKED>(rb, h, acc);
                }
            } else {
#pragma unroll 1
                for (int yy = FP - 1; yy < KS; ++yy)
                    fwd_row_tma<KS, CG, 0, FP, PACKED>(srow + yy * PITCH, vrow + yy * Cfg::VROW, h, acc);
            }
            static_for<0, FP - 1>([&](auto E) {
                constexpr int yy = KS + decltype(E)::value;
                fwd_row_tma<KS, CG, decltype(E)::value + 1, FP, PACKED>(srow + yy * PITCH, vrow + yy * Cfg::VROW, h, acc);
            });

#pragma unroll
            for (int c = 0; c < CG; ++c)
#pragma unroll
                for (int r = 0; r < FP; ++r) {
                    float a = acc[c][r];
                    a += __shfl_xor_sync(0xffffffffu, a, 8);
                    a += __shfl_xor_sync(0xffffffffu, a, 16);
                    res[s][c][r] = a;
                }
            LAB_T(3);  // sweep + reduce
            __syncthreads();  // everyone is done with vs / is before the next refill
            LAB_T(4);  // barrier wait (slowest warp)
        }

#pragma unroll
        for (int c = 0; c < CG; ++c)
#pragma unroll
            for (int r = 0; r < FP; ++r) {
                if ((r & 3) == ch && px_ok) {
                    const long o = ((long)(b * p.C + c0 + c) * Ho + y0 + r) * Wo + px;
                    if (DUAL) {
                        if (p.out[0]) p.out[0][o] = res[0][c][r];
                        if (p.out[1]) p.out[1][o] = res[DUAL ? 1 : 0][c][r];
                        p.blend[o] = p.a * res[0][c][r] + p.b * res[DUAL ? 1 : 0][c][r];
                    } else {
                        p.out[0][o] = res[0][c][r];
                    }
                }
            }
    }
    LAB_T(5);  // stores
    }  // persistent tile loop
}

}  // namespace tai
