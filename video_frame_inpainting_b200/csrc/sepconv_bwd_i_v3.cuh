// Gradient of the separable convolution w.r.t. its (padded) input, for sm_100a
// (persistent, TMA-fed, scatter form; compile-time ks).
//
//   gI[b,c,Y+i,X+j] += gO[b,c,Y,X] * V[b,i,Y,X] * H[b,j,Y,X]       for every source pixel (Y,X), tap (i,j)
//
// which is kernel.cu:120-162 read from the source side: the reference's bounds test
// X<0 || Y<0 || Y>=Ho || X>=Wo (kernel.cu:150) selects exactly the (source, tap) pairs that exist, so
// looping over existing sources and all taps visits the same set of products.
//
// The reference gathers: one thread per gI element, 3 loads per FMA, V and H re-read ks*ks times.
// Here each source pixel's kernels are read ONCE (same TMA slab scheme as the forward kernel) and the
// sweep is the forward sweep transposed:
//   * a warp owns 8 source columns x 8 source rows; lane = (cx = lane&7, ch = lane>>3) keeps the
//     horizontal taps j == ch (mod 4) of its 8 source pixels in registers;
//   * for destination row yy (relative to the tile), source row r contributes through vertical tap yy-r:
//         t[j] = sum_r (V_{yy-r}(r) * gO_c(r)) * H_j(r)          8 FMAs per tap, V from the slab
//     is this lane's contribution to destination column cx + j of that row;
//   * the 4 x 13 values of a pixel column group are staged in shared memory and summed along the
//     anti-diagonals cx + j = const (8 terms) by the warp itself: 58 destination columns per warp-row;
//   * each warp keeps a private ROLLING window of two 8-row groups x 58 destination columns; whenever the
//     four warps have completed a group (one __syncthreads per 8 destination rows) the CTA merges the four
//     windows of that group (they overlap by 50 columns) and adds the 8 x 82 result into gI with red.global
//     while the sweep carries on into the other buffer (neighbouring tiles overlap by ks-1 rows/columns, so
//     gI is zeroed by the launcher first).  A full 58 x 58 window per warp (the first version) cost 54 KB of
//     shared memory and held the kernel at two CTAs per SM; the rolling window needs 15 KB: three CTAs.
#pragma once

#include "common.cuh"
#include "sepconv_common.cuh"
#include "sepconv_bwd_vh_v3.cuh"  // BwdParams
#include "tma.cuh"

namespace tai {

template <int KS>
struct GiV3Cfg {
    static constexpr int J = (KS + 3) / 4;
    static constexpr int WX = 4;
    static constexpr int NT = 32 * WX;
    static constexpr int TILE_W = WX * FNX, TILE_H = FP;
    static constexpr int NCHUNK = 3;
    static constexpr int CH_TAPS = (KS + NCHUNK - 1) / NCHUNK;
    static constexpr int VROW = TILE_H * TILE_W;
    static constexpr int SLAB_FLOATS = NCHUNK * CH_TAPS * VROW;
    static constexpr int DROWS = TILE_H + KS - 1;    // destination rows per tile
    static constexpr int WCOLS = FNX + KS - 1;       // destination columns per warp
    static constexpr int DCOLS = TILE_W + KS - 1;    // destination columns per CTA
    static constexpr int GROWS = 8;                  // destination rows per flush group
    static constexpr int NGROUPS = (DROWS + GROWS - 1) / GROWS;
    static constexpr int WIN_FLOATS = 2 * GROWS * WCOLS;  // per warp: two groups (one being written, one being flushed)
    static constexpr int TS_PITCH = FNX + 1;
    static constexpr int TS_PAD = FNX - 1;                               // zero rows j = -7..-1 and KS..KS+6
    static constexpr int TS_FLOATS = (KS + 2 * TS_PAD) * TS_PITCH + 8;   // so that the diagonal reads need no predicates
    static constexpr int NBAR = 1 + NCHUNK;
    static constexpr size_t smem_bytes() { return (size_t)(SLAB_FLOATS + WX * (WIN_FLOATS + TS_FLOATS)) * 4 + 8 * NBAR; }
};

struct GiV3Maps {
    CUtensorMap h;  // box {32, 8, KS, 1}
    CUtensorMap v;  // box {32, 8, CH_TAPS, 1}
};

// Anti-diagonal sums of one staged row: destination column d receives ts[d - c][c], c = 0..7.
// Split in two so that the 16 shared-memory loads are in flight underneath the next row's FMAs.
template <int KS>
struct GiDiag {
    float a[2][FNX];
};
template <int KS>
__device__ __forceinline__ void gi_diag_load(const float *ts, int lane, GiDiag<KS> &g)
{
    using Cfg = GiV3Cfg<KS>;
    // ts points at row j = 0; rows -7..-1 and KS..KS+6 are zero.  Lanes whose second column does not exist
    // (lane + 32 >= WCOLS) read their first column twice; the result is not stored.
    const int d0 = (lane < Cfg::WCOLS) ? lane : 0, d1 = (lane + 32 < Cfg::WCOLS) ? lane + 32 : d0;
    const float *p0 = ts + d0 * Cfg::TS_PITCH, *p1 = ts + d1 * Cfg::TS_PITCH;
#pragma unroll
    for (int c = 0; c < FNX; ++c) {
        g.a[0][c] = p0[c * (1 - Cfg::TS_PITCH)];
        g.a[1][c] = p1[c * (1 - Cfg::TS_PITCH)];
    }
}
template <int KS>
__device__ __forceinline__ void gi_diag_store(const GiDiag<KS> &g, float *wrow, int lane)
{
    using Cfg = GiV3Cfg<KS>;
    float s0 = (g.a[0][0] + g.a[0][1]) + (g.a[0][2] + g.a[0][3]);
    float s1 = (g.a[1][0] + g.a[1][1]) + (g.a[1][2] + g.a[1][3]);
    s0 += (g.a[0][4] + g.a[0][5]) + (g.a[0][6] + g.a[0][7]);
    s1 += (g.a[1][4] + g.a[1][5]) + (g.a[1][6] + g.a[1][7]);
    if (lane < Cfg::WCOLS) wrow[lane] = s0;
    if (lane + 32 < Cfg::WCOLS) wrow[lane + 32] = s1;
}

// One destination row yy: source rows [RLO, RHI) of this thread reach it (0 <= yy - r < KS).
// Software pipeline over rows: while the 8 x J FMAs of row yy run, the staged values of row yy-1 are on
// their way from shared memory; they are summed and written to the window after the FMAs, then (all lanes
// have consumed the staging buffer: first __syncwarp) row yy is staged (second __syncwarp: visible).
template <int KS, int RLO, int RHI, bool HAS_PREV>
__device__ __forceinline__ void gi_row_v3(const float *__restrict__ vrow, const float (&h)[FP][(KS + 3) / 4],
                                          const float (&go)[FP], float *ts, float *wrow_prev, int cx, int ch, int lane)
{
    using Cfg = GiV3Cfg<KS>;
    constexpr int J = Cfg::J;
    float vg[FP];
#pragma unroll
    for (int r = RLO; r < RHI; ++r) vg[r] = vrow[r * (Cfg::TILE_W - Cfg::VROW)] * go[r];  // tap yy-r of source row r
    GiDiag<KS> g;
    if (HAS_PREV) gi_diag_load<KS>(ts, lane, g);
    float t[J];
#pragma unroll
    for (int jj = 0; jj < J; ++jj) t[jj] = vg[RLO] * h[RLO][jj];
#pragma unroll
    for (int r = RLO + 1; r < RHI; ++r)
#pragma unroll
        for (int jj = 0; jj < J; ++jj) t[jj] = fmaf(vg[r], h[r][jj], t[jj]);
    if (HAS_PREV) {
        gi_diag_store<KS>(g, wrow_prev, lane);
        __syncwarp();
    }
    // stage: ts[j][cx]
#pragma unroll
    for (int jj = 0; jj < J; ++jj)
        if (ch + 4 * jj < KS) ts[(ch + 4 * jj) * Cfg::TS_PITCH + cx] = t[jj];
    __syncwarp();
}

template <int KS>
__global__ void __launch_bounds__(128, 3)
sepconv_bwd_i_v3_kernel(const __grid_constant__ GiV3Maps maps, const BwdParams p)
{
    using Cfg = GiV3Cfg<KS>;
    constexpr int J = Cfg::J, TILE_W = Cfg::TILE_W, TILE_H = Cfg::TILE_H;
    static_assert(Cfg::WCOLS <= 64, "two destination columns per lane");
    extern __shared__ __align__(128) float smem[];
    float *slab = smem;
    float *win = smem + Cfg::SLAB_FLOATS;                       // [4 warps][2 groups][GROWS][WCOLS]
    float *tsb = win + Cfg::WX * Cfg::WIN_FLOATS;               // [4 warps][KS][TS_PITCH]
    uint64_t *bars = reinterpret_cast<uint64_t *>(tsb + Cfg::WX * Cfg::TS_FLOATS);

    const int Ho = p.Ho, Wo = p.Wo;
    const int Hi = Ho + KS - 1, Wi = Wo + KS - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cx = lane & 7, ch = lane >> 3;
    const int ntiles = p.B * p.nty * p.ntx;
    float *ts = tsb + warp * Cfg::TS_FLOATS + Cfg::TS_PAD * Cfg::TS_PITCH;  // row j = 0
    float *mywin = win + warp * Cfg::WIN_FLOATS;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < Cfg::NBAR; ++i) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < Cfg::WX * Cfg::TS_FLOATS; i += Cfg::NT) tsb[i] = 0.f;
    __syncthreads();
    uint32_t parity = 0;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int t = tile;
        const int tx = t % p.ntx;
        t /= p.ntx;
        const int ty = t % p.nty;
        const int b = t / p.nty;
        const int x0 = tx * TILE_W, y0 = ty * TILE_H;  // NOT shifted: a scatter must not visit a source twice
        const int px = x0 + warp * FNX + cx;

        if (threadIdx.x == 0) {
            fence_proxy_async();
            mbar_expect_tx(&bars[0], KS * Cfg::VROW * 4);
            tma_load_4d(slab, &maps.h, &bars[0], x0, y0, 0, b);  // out-of-range rows / columns arrive as zeros
        }
        mbar_wait(&bars[0], parity);
        float h[FP][J];
        {
            const float *hs = slab + ch * Cfg::VROW + warp * FNX + cx;
#pragma unroll
            for (int jj = 0; jj < J; ++jj)
#pragma unroll
                for (int r = 0; r < FP; ++r)
                    h[r][jj] = (ch + 4 * jj < KS) ? hs[(4 * jj) * Cfg::VROW + r * TILE_W] : 0.f;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_proxy_async();
#pragma unroll
            for (int q = 0; q < Cfg::NCHUNK; ++q) {
                mbar_expect_tx(&bars[1 + q], Cfg::CH_TAPS * Cfg::VROW * 4);
                tma_load_4d(slab + q * Cfg::CH_TAPS * Cfg::VROW, &maps.v, &bars[1 + q], x0, y0, q * Cfg::CH_TAPS, b);
            }
            if (tile + (int)gridDim.x < ntiles) {
                int n = tile + gridDim.x;
                const int nx0 = (n % p.ntx) * TILE_W;
                n /= p.ntx;
                const int ny0 = (n % p.nty) * TILE_H, nb = n / p.nty;
                tma_prefetch_l2_4d(&maps.h, nx0, ny0, 0, nb);
#pragma unroll
                for (int q = 0; q < Cfg::NCHUNK; ++q) tma_prefetch_l2_4d(&maps.v, nx0, ny0, q * Cfg::CH_TAPS, nb);
            }
        }
        const float *vrow = slab + warp * FNX + cx;
        auto wrow = [&](int yy) {  // this warp's window row of destination row yy (rolling: group parity, row in group)
            return mywin + (((yy >> 3) & 1) * Cfg::GROWS + (yy & 7)) * Cfg::WCOLS;
        };

        for (int c = 0; c < p.C; ++c) {
            float *gdst = p.gin + ((long)(b * p.C + c) * Hi + y0) * Wi + x0;
            // Merge group g of the four warp windows and add it into gI.  Every thread calls it at the same
            // point of the sweep; the barrier also separates this group's buffer from its reuse two groups later.
            auto flush = [&](int g) {
                __syncthreads();
                const int D = threadIdx.x;
                if (D < Cfg::DCOLS && x0 + D < Wi) {
                    const float *wb = win + (g & 1) * Cfg::GROWS * Cfg::WCOLS;
                    const int nrow = min(Cfg::GROWS, Cfg::DROWS - g * Cfg::GROWS);
                    for (int r = 0; r < nrow; ++r) {
                        float sum = 0.f;
#pragma unroll
                        for (int w = 0; w < Cfg::WX; ++w) {
                            const int dc = D - w * FNX;
                            if (dc >= 0 && dc < Cfg::WCOLS) sum += wb[w * Cfg::WIN_FLOATS + r * Cfg::WCOLS + dc];
                        }
                        const int yy = g * Cfg::GROWS + r;
                        if (y0 + yy < Hi) atomicAdd(gdst + (long)yy * Wi + D, sum);
                    }
                }
            };

            float go[FP];
#pragma unroll
            for (int r = 0; r < FP; ++r)
                go[r] = (px < Wo && y0 + r < Ho) ? __ldg(p.gout + ((long)(b * p.C + c) * Ho + y0 + r) * Wo + px) : 0.f;

            constexpr int PRO_CHUNKS = (FP - 2) / Cfg::CH_TAPS + 1;  // chunks touched by the prologue rows
            if (c == 0) {
#pragma unroll
                for (int q = 0; q < PRO_CHUNKS; ++q) mbar_wait(&bars[1 + q], parity);
            }
            static_for<0, FP - 1>([&](auto YY) {
                constexpr int yy = decltype(YY)::value;
                gi_row_v3<KS, 0, yy + 1, (yy > 0)>(vrow + yy * Cfg::VROW, h, go, ts, wrow(yy - 1), cx, ch, lane);
            });
#pragma unroll
            for (int q = 0; q < Cfg::NCHUNK; ++q) {
                const int lo = max(FP - 1, q * Cfg::CH_TAPS);
                const int hi = (q == Cfg::NCHUNK - 1) ? KS : min(KS, (q + 1) * Cfg::CH_TAPS);
                if (q >= PRO_CHUNKS && c == 0) mbar_wait(&bars[1 + q], parity);
#pragma unroll 1
                for (int yy = lo; yy < hi; ++yy) {
                    gi_row_v3<KS, 0, FP, true>(vrow + yy * Cfg::VROW, h, go, ts, wrow(yy - 1), cx, ch, lane);
                    if ((yy & 7) == 0) flush((yy >> 3) - 1);  // row yy-1, the last of its group, has just been stored
                }
            }
            static_for<0, FP - 1>([&](auto E) {
                constexpr int yy = KS + decltype(E)::value;
                gi_row_v3<KS, decltype(E)::value + 1, FP, true>(vrow + yy * Cfg::VROW, h, go, ts, wrow(yy - 1), cx, ch,
                                                               lane);
                if ((yy & 7) == 0) flush((yy >> 3) - 1);
            });
            {   // drain the pipeline: the last destination row, then the last group
                constexpr int yy = Cfg::DROWS - 1;
                GiDiag<KS> g;
                gi_diag_load<KS>(ts, lane, g);
                gi_diag_store<KS>(g, wrow(yy), lane);
                flush(yy >> 3);
            }
        }
        parity ^= 1;
    }
}

}  // namespace tai
