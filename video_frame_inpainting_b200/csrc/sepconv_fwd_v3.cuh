// Forward separable convolution for sm_100a: persistent, TMA-fed kernel (compile-time ks).
//
//   O[b,c,y,x] = sum_i V[b,i,y,x] * ( sum_j H[b,j,y,x] * I[b,c,y+i,x+j] )          (kernel.cu:19-47)
//
// Work decomposition
//   * CTA = 4 warps, tile = 8 rows x 32 columns; a warp owns 8 columns x 8 rows.  Lane = (cx = lane&7,
//     ch = lane>>3); lane group ch owns the horizontal taps j == ch (mod 4) and keeps them for its
//     8 pixels in registers (8 x ceil(ks/4) = 104 for ks = 51), so every FMA of the row sums reads H from a
//     register and I from the shared-memory halo (one LDS feeds 8 FMAs: the 8 output rows of a thread see
//     one input row at 8 different vertical taps).  The four tap groups are summed once per tile with
//     two shuffles.
//   * Persistent CTAs (3 per SM) walk tiles blockIdx.x, +gridDim.x, ...
// Data movement (the part that decides the speed; V and H are 98 % of the bytes and are read once)
//   * one shared-memory SLAB is time-shared by the two kernel maps of a tile.  Thread 0 issues a single
//     cp.async.bulk.tensor (TMA) for the H box [ks taps][8 rows][32 cols]; when its mbarrier flips every
//     thread copies its 8 x J taps to registers; then the same slab is refilled with the V box in three
//     tap chunks (3 TMAs, 3 mbarriers) so the sweep starts when the first third has landed and the rest
//     streams in underneath the FMA work.  No LDG is issued for V or H: 100+ LDGs per thread throttle on
//     the LSU's outstanding-request limit (measured: the LDG-based load phase took as long as the sweep).
//   * while a tile is being filtered, the slabs of the CTA's next tile are pulled into L2 with
//     cp.async.bulk.prefetch.tensor, so the next TMA fills are L2 hits and HBM streams continuously.
//   * the (8+ks-1) x (32+ks-1) input halo is staged with batched LDG/STS; with PAD the replication pad of
//     tai.py:170-171 is folded in (clamped source coordinates).  DUAL filters both predictions and
//     applies the blend of tai.py:105 / twi.py:105 in the epilogue.
#pragma once

#include "common.cuh"
#include "sepconv_common.cuh"
#include "tma.cuh"

#ifndef TAI_HALO_RB
#define TAI_HALO_RB ((Cfg::ROWS + Cfg::NT / 32 - 1) / (Cfg::NT / 32))
#endif

namespace tai {

#ifdef TAI_LAB_TIMING
__device__ unsigned long long g_lab_phase[8];
#define LAB_T(i)                                                             \
    do {                                                                     \
        if (threadIdx.x == 0) {                                              \
            const long long now_ = clock64();                                \
            atomicAdd(&g_lab_phase[i], (unsigned long long)(now_ - lab_t_)); \
            lab_t_ = now_;                                                   \
        }                                                                    \
    } while (0)
#else
#define LAB_T(i)
#endif

template <int KS>
struct FwdV3Cfg {
    static constexpr int J = (KS + 3) / 4;
    static constexpr int WX = 4;
    static constexpr int NT = 32 * WX;
    static constexpr int TILE_W = WX * FNX, TILE_H = FP;
    static constexpr int PITCH = TILE_W + 4 * J;
    static constexpr int ROWS = TILE_H + KS - 1;
    static constexpr int NCHUNK = 3;
    static constexpr int CH_TAPS = (KS + NCHUNK - 1) / NCHUNK;      // taps per V chunk
    static constexpr int VROW = TILE_H * TILE_W;                    // floats per tap in the slab
    static constexpr int SLAB_FLOATS = NCHUNK * CH_TAPS * VROW;     // >= KS * VROW
    static constexpr int NBAR = 1 + NCHUNK;
    static constexpr size_t smem_bytes(int cg) { return (size_t)(SLAB_FLOATS + cg * ROWS * PITCH) * 4 + 8 * NBAR; }
};

struct FwdV3Maps {
    CUtensorMap h[2];  // box {32, KS, 8, 1}, 128-byte swizzle (make_kernel_map_tmap_swz)
    CUtensorMap v[2];  // box {32, 8, CH_TAPS, 1}
};

template <int KS, int CG, int RLO, int RHI>
__device__ __forceinline__ void fwd_row_v3(const float *__restrict__ srow, const float *__restrict__ vrow,
                                           const float (&h)[FP][(KS + 3) / 4], float (&acc)[CG][FP])
{
    using Cfg = FwdV3Cfg<KS>;
    constexpr int J = Cfg::J;
    constexpr int CSTRIDE = Cfg::ROWS * Cfg::PITCH;
    float v[FP];
#pragma unroll
    for (int r = RLO; r < RHI; ++r) v[r] = vrow[r * (Cfg::TILE_W - Cfg::VROW)];  // tap yy-r, output row r
#pragma unroll
    for (int c = 0; c < CG; ++c) {
        float iv[J];
#pragma unroll
        for (int jj = 0; jj < J; ++jj) iv[jj] = srow[c * CSTRIDE + 4 * jj];
        // tap-outer order: the row sums are independent FMA chains that interleave (row-outer order makes ptxas
        // emit one serial chain after the other: 4-cycle dependent issue with 3 warps per scheduler)
        float s[FP];
#pragma unroll
        for (int r = RLO; r < RHI; ++r) s[r] = h[r][0] * iv[0];
#pragma unroll
        for (int jj = 1; jj < J; ++jj)
#pragma unroll
            for (int r = RLO; r < RHI; ++r) s[r] = fmaf(h[r][jj], iv[jj], s[r]);
#pragma unroll
        for (int r = RLO; r < RHI; ++r) acc[c][r] = fmaf(v[r], s[r], acc[c][r]);
    }
}

template <int KS, int CG, bool PAD, bool DUAL>
// small windows (ks <= 16) are HBM-bound and keep few taps in registers: six CTAs per SM put more boxes in flight
// (ks = 13, C = 1, [64,1,512,512]: 65.5 -> 68.9 % of the measured copy bandwidth; at C = 3 four CTAs were no gain)
__global__ void __launch_bounds__(128, (CG == 1 ? (KS <= 16 ? 6 : KS <= 28 ? 4 : TAI_FWD_MIN_CTAS) : 2))
sepconv_fwd_v3_kernel(const __grid_constant__ FwdV3Maps maps, const FwdParams p)
{
    using Cfg = FwdV3Cfg<KS>;
    constexpr int J = Cfg::J, PITCH = Cfg::PITCH, ROWS = Cfg::ROWS, TILE_W = Cfg::TILE_W, TILE_H = Cfg::TILE_H;
    constexpr int CSTRIDE = ROWS * PITCH;
    constexpr int NS = DUAL ? 2 : 1;
    extern __shared__ __align__(1024) float smem[];
    float *slab = smem;                       // H: [TILE_H][taps][TILE_W] swizzled, V: [taps][TILE_H][TILE_W]; written by TMA only
    float *is = smem + Cfg::SLAB_FLOATS;      // [CG][ROWS][PITCH]
    uint64_t *bars = reinterpret_cast<uint64_t *>(is + CG * CSTRIDE);  // [0]: H box, [1..3]: V chunks

    const int Ho = p.Ho, Wo = p.Wo;
    const int Hi = Ho + KS - 1, Wi = Wo + KS - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cx = lane & 7, ch = lane >> 3;
    const long plane = (long)Ho * Wo;
    const int ntiles = p.B * p.nty * p.ntx;

    if (threadIdx.x == 0) {
        if ((__cvta_generic_to_shared(slab) & 1023) != 0) __trap();  // the swizzle pattern is tied to 1024-byte blocks
#pragma unroll
        for (int i = 0; i < Cfg::NBAR; ++i) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t parity = 0;  // every barrier completes exactly once per (tile, stream, channel group)
    int swz[8];
    swz_table(warp * FNX + cx, ch, swz);
#ifdef TAI_LAB_TIMING
    long long lab_t_ = clock64();
#endif

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int t = tile;
        const int tx = t % p.ntx;
        t /= p.ntx;
        const int ty = t % p.nty;
        const int b = t / p.nty;
        // Tiles that would stick out are shifted back inside (their overlap recomputes identical values).
        const int x0 = max(0, min(tx * TILE_W, Wo - TILE_W));
        const int y0 = min(ty * TILE_H, Ho - TILE_H);  // host guarantees Ho >= TILE_H
        const int px_raw = x0 + warp * FNX + cx;
        const bool px_ok = px_raw < Wo;

        float res[NS][CG][FP];

        for (int c0 = 0; c0 < p.C; c0 += CG) {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                // ---- H box -> slab (TMA); halo -> smem (cp.async) ----
                if (threadIdx.x == 0) {
                    fence_proxy_async();  // generic-proxy reads of the slab (previous sweep) precede the refill
                    mbar_expect_tx(&bars[0], KS * Cfg::VROW * 4);
                    tma_load_4d(slab, &maps.h[s], &bars[0], x0, 0, y0, b);
                }
                {
                    const float *__restrict__ in = p.in[s];
                    // per-lane column bookkeeping is row independent
                    int coff[(PITCH + 31) / 32];
                    bool cok[(PITCH + 31) / 32];
#pragma unroll
                    for (int k = 0; k < (PITCH + 31) / 32; ++k) {
                        const int rx = lane + 32 * k, gx = x0 + rx;
                        if (PAD) {
                            cok[k] = rx < TILE_W + KS - 1;
                            coff[k] = clampi(gx - KS / 2, 0, Wo - 1);
                        } else {
                            cok[k] = rx < TILE_W + KS - 1 && gx < Wi;
                            coff[k] = cok[k] ? gx : 0;
                        }
                    }
                    // LDG -> STS in register batches (an LDGSTS costs ~8 LSU cycles, LDG + STS ~3)
                    constexpr int NK = (PITCH + 31) / 32;
                    // rows per batch and warp.  Single stream: ALL of a warp's rows in one batch (the H registers are not
                    // live yet, so 45 staging registers are free): one memory round trip per tile instead of four
                    // (0.406 -> 0.393 ms at B = 160).  The dual-stream kernel takes two batches of 8 rows: with the first
                    // stream's results live, one batch of 15 costs more than it hides (0.764 -> 0.836 ms, measured),
                    // 4-row batches (four round trips) are 1.3 % behind (0.764 vs 0.754 ms).
                    constexpr int RB = DUAL ? 8 : TAI_HALO_RB;
                    for (int c = 0; c < CG; ++c) {
                        const float *src = PAD ? in + ((long)(b * p.C + c0 + c)) * plane
                                               : in + ((long)(b * p.C + c0 + c)) * Hi * Wi;
                        for (int ry0 = warp * RB; ry0 < ROWS; ry0 += (Cfg::NT / 32) * RB) {
                            float tmp[RB][NK];
#pragma unroll
                            for (int q = 0; q < RB; ++q) {
                                const int gy = y0 + min(ry0 + q, ROWS - 1);
                                const float *grow = PAD ? src + (long)clampi(gy - KS / 2, 0, Ho - 1) * Wo : src + (long)gy * Wi;
#pragma unroll
                                for (int k = 0; k < NK; ++k) tmp[q][k] = cok[k] ? __ldg(grow + coff[k]) : 0.f;
                            }
#pragma unroll
                            for (int q = 0; q < RB; ++q) {
                                if (ry0 + q < ROWS) {
                                    float *drow = is + c * CSTRIDE + (ry0 + q) * PITCH + lane;
#pragma unroll
                                    for (int k = 0; k < NK; ++k)
                                        if (lane + 32 * k < PITCH) drow[32 * k] = tmp[q][k];
                                }
                            }
                        }
                    }
                }
                LAB_T(0);
                // ---- this lane's horizontal taps: slab -> registers ----
                mbar_wait(&bars[0], parity);
                LAB_T(1);
                float h[FP][J];
                // The H image is [row][tap][col]: rows 0 .. R0-1 cover the slab bytes of V chunk 0.  They are copied
                // first, so that the first V chunk can be requested while the other rows are still being copied
                // (its TMA latency used to be a serial phase of every tile: "wait for the first V chunk").
                constexpr int R0 = (Cfg::CH_TAPS * Cfg::VROW + KS * 32 - 1) / (KS * 32);   // H rows overlapping V chunk 0
                static_assert(R0 < FP, "the first V chunk must leave H rows to copy under its load");
#pragma unroll
                for (int jj = 0; jj < J; ++jj)
#pragma unroll
                    for (int r = 0; r < R0; ++r)
                        h[r][jj] = (ch + 4 * jj < KS) ? slab[(r * KS + 4 * jj) * 32 + swz[(r * KS + 4 * jj) & 7]] : 0.f;
                __syncthreads();  // the first R0 rows of H are in registers everywhere
                if (threadIdx.x == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(&bars[1], Cfg::CH_TAPS * Cfg::VROW * 4);
                    tma_load_4d(slab, &maps.v[s], &bars[1], x0, y0, 0, b);
                }
#pragma unroll
                for (int jj = 0; jj < J; ++jj)
#pragma unroll
                    for (int r = R0; r < FP; ++r)
                        h[r][jj] = (ch + 4 * jj < KS) ? slab[(r * KS + 4 * jj) * 32 + swz[(r * KS + 4 * jj) & 7]] : 0.f;
                __syncthreads();  // H is in registers everywhere; the halo is complete
                LAB_T(2);
                // ---- the other V chunks; prefetch the next tile's boxes into L2 ----
                if (threadIdx.x == 0) {
                    fence_proxy_async();
#pragma unroll
                    for (int q = 1; q < Cfg::NCHUNK; ++q) {
                        mbar_expect_tx(&bars[1 + q], Cfg::CH_TAPS * Cfg::VROW * 4);
                        tma_load_4d(slab + q * Cfg::CH_TAPS * Cfg::VROW, &maps.v[s], &bars[1 + q], x0, y0,
                                    q * Cfg::CH_TAPS, b);
                    }
                    // L2 prefetch of the boxes of the NEXT item only (the second stream of this tile, or the first
                    // stream of the CTA's next tile).  Prefetching both streams of the next tile a whole item
                    // early put 8 boxes (209 KB) per CTA in flight: 93 MB for 444 CTAs, which the 126 MB L2 evicted
                    // before use -- ncu r01 read 3.24 GB from DRAM for 2.19 GB of maps.
                    if (DUAL && s == 0) {
                        tma_prefetch_l2_4d(&maps.h[1], x0, 0, y0, b);
#pragma unroll
                        for (int q = 0; q < Cfg::NCHUNK; ++q) tma_prefetch_l2_4d(&maps.v[1], x0, y0, q * Cfg::CH_TAPS, b);
                    } else if (c0 + CG >= p.C && tile + (int)gridDim.x < ntiles) {
                        int n = tile + gridDim.x;
                        const int ntx_ = n % p.ntx;
                        n /= p.ntx;
                        const int nty_ = n % p.nty;
                        const int nb = n / p.nty;
                        const int nx0 = max(0, min(ntx_ * TILE_W, Wo - TILE_W)), ny0 = min(nty_ * TILE_H, Ho - TILE_H);
                        tma_prefetch_l2_4d(&maps.h[0], nx0, 0, ny0, nb);
#pragma unroll
                        for (int q = 0; q < Cfg::NCHUNK; ++q) tma_prefetch_l2_4d(&maps.v[0], nx0, ny0, q * Cfg::CH_TAPS, nb);
                    }
                }

                float acc[CG][FP];
#pragma unroll
                for (int c = 0; c < CG; ++c)
#pragma unroll
                    for (int r = 0; r < FP; ++r) acc[c][r] = 0.f;

                const float *srow = is + warp * FNX + cx + ch;
                const float *vrow = slab + warp * FNX + cx;  // tap 0, row 0

                // the prologue rows touch vertical taps 0..FP-2: wait for every chunk that holds one of them
                constexpr int PRO_CHUNKS = (FP - 2) / Cfg::CH_TAPS + 1;
#pragma unroll
                for (int q = 0; q < PRO_CHUNKS; ++q) mbar_wait(&bars[1 + q], parity);
                LAB_T(3);
                // prologue: input rows 0..6 (output rows 0..yy are inside the window)
                static_for<0, FP - 1>([&](auto YY) {
                    constexpr int yy = decltype(YY)::value;
                    fwd_row_v3<KS, CG, 0, yy + 1>(srow + yy * PITCH, vrow + yy * Cfg::VROW, h, acc);
                });
                // steady state, one V chunk at a time (row yy needs vertical taps <= yy)
#pragma unroll
                for (int q = 0; q < Cfg::NCHUNK; ++q) {
                    const int lo = max(FP - 1, q * Cfg::CH_TAPS);
                    const int hi = (q == Cfg::NCHUNK - 1) ? KS : min(KS, (q + 1) * Cfg::CH_TAPS);
                    if (q >= PRO_CHUNKS) mbar_wait(&bars[1 + q], parity);
#pragma unroll 1
                    for (int yy = lo; yy < hi; ++yy)
                        fwd_row_v3<KS, CG, 0, FP>(srow + yy * PITCH, vrow + yy * Cfg::VROW, h, acc);
                }
                // epilogue: input rows ks..ks+6
                static_for<0, FP - 1>([&](auto E) {
                    constexpr int yy = KS + decltype(E)::value;
                    fwd_row_v3<KS, CG, decltype(E)::value + 1, FP>(srow + yy * PITCH, vrow + yy * Cfg::VROW, h, acc);
                });

                // ---- sum the four tap groups ----
#pragma unroll
                for (int c = 0; c < CG; ++c)
#pragma unroll
                    for (int r = 0; r < FP; ++r) {
                        float a = acc[c][r];
                        a += __shfl_xor_sync(0xffffffffu, a, 8);
                        a += __shfl_xor_sync(0xffffffffu, a, 16);
                        res[s][c][r] = a;
                    }
                parity ^= 1;
                LAB_T(4);
                __syncthreads();  // every warp is done with the slab and the halo before the next refill
                LAB_T(5);
            }

            // ---- stores (+ blend): lane group ch writes rows r == ch (mod 4) ----
#pragma unroll
            for (int c = 0; c < CG; ++c)
#pragma unroll
                for (int r = 0; r < FP; ++r) {
                    if ((r & 3) == ch && px_ok) {
                        const long o = ((long)(b * p.C + c0 + c) * Ho + y0 + r) * Wo + px_raw;
                        if (DUAL) {
                            if (p.out[0]) p.out[0][o] = res[0][c][r];
                            if (p.out[1]) p.out[1][o] = res[NS - 1][c][r];
                            p.blend[o] = p.a * res[0][c][r] + p.b * res[NS - 1][c][r];
                        } else {
                            p.out[0][o] = res[0][c][r];
                        }
                    }
                }
            LAB_T(6);
        }
    }
}

}  // namespace tai
