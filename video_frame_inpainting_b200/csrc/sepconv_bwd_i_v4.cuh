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
//         t[k] = sum_r (V_{yy-r}(r) * gO_c(r)) * H_{ch+4k}(r)          8 FMAs per tap, V from the slab
//     is this lane's contribution to destination column cx + ch + 4k of that row;
//   * the 32 x 13 values of a warp-row have to be summed into 58 destination columns (8 terms each).  This
//     is where the first version of this kernel spent its time (13 STS with 2-way bank conflicts + 16 LDS
//     per lane and row: the kernel ran at 76 % of the shared-memory wavefront limit, ncu r01).  Now:
//       1. lanes cx and cx+4 (same tap group) exchange their odd tap slots with one SHFL each and add what
//          they receive onto their even slots: tap ch+4k of column cx+4 and tap ch+4(k+1) of column cx
//          land on the same destination column.  The lanes cx >= 4 keep their even taps two physical slots
//          higher (decided when H is copied to registers), so both directions use the same register names:
//          6 SHFL + 6 FADD, no select, and 7 values per lane are left, 8 destination columns apart;
//       2. those are staged as ts[4 * dcol + ch]: the four contributors of a destination column are
//          adjacent, the 32 lanes of a store hit 32 different banks (4cx + 5ch mod 32 is a bijection),
//          and a destination column is read back with ONE 128-bit load: 7 STS + 2 LDS.128 per lane and row;
//   * each warp keeps a private ROLLING window of two 8-row groups x 58 destination columns; whenever the
//     four warps have completed a group (one __syncthreads per 8 destination rows) the CTA merges the four
//     windows of that group (they overlap by 50 columns) and adds the 8 x 82 result into gI with red.global
//     while the sweep carries on into the other buffer (neighbouring tiles overlap by ks-1 rows/columns, so
//     gI is zeroed by the launcher first).
//   * FOLD (C == 1): gO is multiplied into the H registers once per tile instead of into V at every row.
#pragma once

#include "common.cuh"
#include "sepconv_common.cuh"
#include "sepconv_bwd_vh_v3.cuh"  // BwdParams
#include "tma.cuh"

namespace tai {

template <int KS>
struct GiV4Cfg {
    static constexpr int J = (KS + 3) / 4;
    static constexpr int JP = J | 1;             // physical tap slots per lane (odd; a padding slot holds a zero tap)
    static constexpr int E = (JP + 1) / 2;       // values per lane after the pair exchange
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
    static constexpr int WPITCH = (WCOLS + 3) / 4 * 4;    // window row pitch: whole 16-byte column quads (the pad stays zero)
    static constexpr int WQUADS = WPITCH / 4;             // column quads per warp window row
    static constexpr int DQUADS = (DCOLS + 3) / 4;        // column quads per CTA row
    static constexpr int WIN_FLOATS = 2 * GROWS * WPITCH; // per warp: two groups (one being written, one being flushed)
    static constexpr int TS_FLOATS = 4 * 64;              // per warp: [destination column 0..63][contributor ch]
    static constexpr int NBAR = 1 + NCHUNK;
    static constexpr size_t smem_bytes() { return (size_t)(SLAB_FLOATS + WX * (WIN_FLOATS + TS_FLOATS)) * 4 + 8 * NBAR; }
    static_assert(FNX == 8 && FP == 8, "lane layout: 8 source columns x 8 source rows per warp");
    static_assert(WCOLS <= 64, "two destination columns per lane");
    static_assert(FNX + 2 + 8 * (E - 1) < 64, "staging buffer holds destination columns 0..63");
    static_assert((SLAB_FLOATS % 4) == 0 && (WIN_FLOATS % 4) == 0, "16-byte alignment of the staging buffer");
    static_assert(FNX % 4 == 0, "a column quad of the CTA row is a column quad of every warp window");
};

struct GiV4Maps {
    CUtensorMap h;  // box {32, KS, 8, 1}, 128-byte swizzle (make_kernel_map_tmap_swz)
    CUtensorMap v;  // box {32, 8, CH_TAPS, 1}
};

// Two adjacent floats added to global memory with one reduction (sm_90+; 8-byte aligned address).  The flush is bound
// by the SM's rate of global reductions per lane, not by bytes: pairs halve it.
__device__ __forceinline__ void red_add_v2(float *addr, float a, float b)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// The staged row: destination column d of the warp is the sum of the four floats at ts[4d..4d+3].
struct GiV4Diag {
    float4 a, b;  // destination columns lane and lane + 32
};
__device__ __forceinline__ void gi4_diag_load(const float *ts, int lane, GiV4Diag &g)
{
    g.a = reinterpret_cast<const float4 *>(ts)[lane];
    g.b = reinterpret_cast<const float4 *>(ts)[lane + 32];
}
template <int KS>
__device__ __forceinline__ void gi4_diag_store(const GiV4Diag &g, float *wrow, int lane)
{
    using Cfg = GiV4Cfg<KS>;
    const float s0 = (g.a.x + g.a.y) + (g.a.z + g.a.w);
    const float s1 = (g.b.x + g.b.y) + (g.b.z + g.b.w);
    if (lane < Cfg::WCOLS) wrow[lane] = s0;
    if (lane + 32 < Cfg::WCOLS) wrow[lane + 32] = s1;
}

// One destination row yy: source rows [RLO, RHI) of this thread reach it (0 <= yy - r < KS).
// Software pipeline over rows: while the 8 x JP FMAs of row yy run, the staged sums of row yy-1 are on
// their way from shared memory; they are added up and written to the window after the FMAs, then (all
// lanes have consumed the staging buffer: first __syncwarp) row yy is staged (second __syncwarp: visible).
template <int KS, int RLO, int RHI, bool HAS_PREV, bool FOLD>
__device__ __forceinline__ void gi_row_v4(const float *__restrict__ vrow,
                                          const float (&h)[FP][GiV4Cfg<KS>::JP], const float (&go)[FP],
                                          const float *ts, float *st0, float *stq, float *wrow_prev, int lane)
{
    using Cfg = GiV4Cfg<KS>;
    constexpr int JP = Cfg::JP, E = Cfg::E;
    float vg[FP];
#pragma unroll
    for (int r = RLO; r < RHI; ++r) {
        const float v = vrow[r * (Cfg::TILE_W - Cfg::VROW)];  // tap yy-r of source row r
        vg[r] = FOLD ? v : v * go[r];
    }
    GiV4Diag g;
    if (HAS_PREV) gi4_diag_load(ts, lane, g);
    float t[JP];
#pragma unroll
    for (int s = 0; s < JP; ++s) t[s] = vg[RLO] * h[RLO][s];
#pragma unroll
    for (int r = RLO + 1; r < RHI; ++r)
#pragma unroll
        for (int s = 0; s < JP; ++s) t[s] = fmaf(vg[r], h[r][s], t[s]);
    // pair exchange (cx, cx+4): odd slots travel, even slots collect
    float red[E];
    red[0] = t[0];
#pragma unroll
    for (int q = 1; q < E; ++q) red[q] = t[2 * q] + __shfl_xor_sync(0xffffffffu, t[2 * q - 1], 4);
    if (HAS_PREV) {
        gi4_diag_store<KS>(g, wrow_prev, lane);
        __syncwarp();
    }
    st0[0] = red[0];
#pragma unroll
    for (int q = 1; q < E; ++q) stq[32 * q] = red[q];
    __syncwarp();
}

template <int KS, bool FOLD>
// small windows keep few taps in registers: six CTAs per SM for ks <= 16 (80 registers; ks = 13 backward + 10 %), four for
// ks <= 28 (ks = 25: + 4 %) put more boxes in flight where the op is HBM-bound
__global__ void __launch_bounds__(128, (KS <= 16 ? 6 : KS <= 28 ? 4 : 3))
sepconv_bwd_i_v4_kernel(const __grid_constant__ GiV4Maps maps, const BwdParams p)
{
    using Cfg = GiV4Cfg<KS>;
    constexpr int JP = Cfg::JP, E = Cfg::E, TILE_W = Cfg::TILE_W, TILE_H = Cfg::TILE_H;
    extern __shared__ __align__(1024) float smem[];
    float *slab = smem;
    float *win = smem + Cfg::SLAB_FLOATS;                       // [4 warps][2 groups][GROWS][WCOLS]
    float *tsb = win + Cfg::WX * Cfg::WIN_FLOATS;               // [4 warps][64][4]
    uint64_t *bars = reinterpret_cast<uint64_t *>(tsb + Cfg::WX * Cfg::TS_FLOATS);

    const int Ho = p.Ho, Wo = p.Wo;
    const int Hi = Ho + KS - 1, Wi = Wo + KS - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cx = lane & 7, ch = lane >> 3;
    const bool hi = cx >= 4;
    const int ntiles = p.B * p.nty * p.ntx;
    const bool pairs = (Wi % 2 == 0) && ((reinterpret_cast<uintptr_t>(p.gin) & 7) == 0);   // 64-bit reductions possible
    const float *ts = tsb + warp * Cfg::TS_FLOATS;
    // after the pair exchange value q of a lane belongs to destination column cx + ch + 8q (lanes cx < 4) or
    // cx + ch + 8((q-1) mod E) (lanes cx >= 4); contributor slot = ch
    float *st_base = tsb + warp * Cfg::TS_FLOATS + 4 * (cx + ch) + ch;
    float *stq = st_base - (hi ? 32 : 0);
    float *st0 = st_base + (hi ? 32 * (E - 1) : 0);
    float *mywin = win + warp * Cfg::WIN_FLOATS;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < Cfg::NBAR; ++i) mbar_init(&bars[i], 1);
        mbar_fence_init();
        if ((__cvta_generic_to_shared(slab) & 1023) != 0) __trap();  // the swizzle pattern is tied to 1024-byte blocks
    }
    for (int i = threadIdx.x; i < Cfg::WX * Cfg::TS_FLOATS; i += Cfg::NT) tsb[i] = 0.f;  // unwritten slots stay zero
    for (int i = threadIdx.x; i < Cfg::WX * Cfg::WIN_FLOATS; i += Cfg::NT) win[i] = 0.f;  // so do the pad columns of the windows
    int swz[8];
    swz_table(warp * FNX + cx, ch, swz);
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
            tma_load_4d(slab, &maps.h, &bars[0], x0, 0, y0, b);  // out-of-range rows / columns arrive as zeros
        }
        float go[FP];
        if (FOLD) {
#pragma unroll
            for (int r = 0; r < FP; ++r)
                go[r] = (px < Wo && y0 + r < Ho) ? __ldg(p.gout + ((long)b * Ho + y0 + r) * Wo + px) : 0.f;
        }
        mbar_wait(&bars[0], parity);
        // physical slot s holds tap ch + 4k: k = s for the lanes cx < 4; the lanes cx >= 4 keep odd slots in
        // place and rotate the even ones up by two (see the pair exchange in gi_row_v4)
        float h[FP][JP];
        {
#pragma unroll
            for (int s = 0; s < JP; ++s) {
                const int kh = (s & 1) ? s : (s == 0 ? JP - 1 : s - 2);
                const int k = hi ? kh : s;     // k and s have the same parity: the swizzle slot is a compile-time index
                const bool ok = ch + 4 * k < KS;
                const float *hk = slab + (4 * k) * 32;
#pragma unroll
                for (int r = 0; r < FP; ++r) {
                    const float hv = ok ? hk[r * KS * 32 + swz[(r * KS + 4 * s) & 7]] : 0.f;
                    h[r][s] = FOLD ? hv * go[r] : hv;
                }
            }
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
                tma_prefetch_l2_4d(&maps.h, nx0, 0, ny0, nb);
#pragma unroll
                for (int q = 0; q < Cfg::NCHUNK; ++q) tma_prefetch_l2_4d(&maps.v, nx0, ny0, q * Cfg::CH_TAPS, nb);
            }
        }
        const float *vrow = slab + warp * FNX + cx;
        auto wrow = [&](int yy) {  // this warp's window row of destination row yy (rolling: group parity, row in group)
            return mywin + (((yy >> 3) & 1) * Cfg::GROWS + (yy & 7)) * Cfg::WPITCH;
        };

        const int nch = FOLD ? 1 : p.C;
        for (int c = 0; c < nch; ++c) {
            float *gdst = p.gin + ((long)(b * p.C + c) * Hi + y0) * Wi + x0;
            // Merge group g of the four warp windows and add it into gI.  Every thread calls it at the same
            // point of the sweep; the barrier also separates this group's buffer from its reuse two groups later.
            auto flush = [&](int g) {
                __syncthreads();
                // All 128 threads: work item = (row of the group, quad of 4 destination columns).  Warp w's window
                // holds destination columns 8w .. 8w+57 at index D - 8w, so a CTA quad Q is quad Q - 2w of window w:
                // one 128-bit load per contributing warp, no per-element index arithmetic.  (First version: a thread
                // per destination column, 8 rows x up to 4 scalar loads each, 82 of 128 threads busy: the flush was
                // 17 % of the kernel in an ablation build.)
                const float *wb = win + (g & 1) * Cfg::GROWS * Cfg::WPITCH;
                const int nrow = min(Cfg::GROWS, Cfg::DROWS - g * Cfg::GROWS);
                for (int item = threadIdx.x; item < nrow * Cfg::DQUADS; item += Cfg::NT) {
                    const int r = item / Cfg::DQUADS, Q = item - r * Cfg::DQUADS;
                    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int w = 0; w < Cfg::WX; ++w) {
                        const int q = Q - w * (FNX / 4);
                        if (q >= 0 && q < Cfg::WQUADS) {
                            const float4 v = *reinterpret_cast<const float4 *>(wb + w * Cfg::WIN_FLOATS + r * Cfg::WPITCH + 4 * q);
                            sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
                        }
                    }
                    const int yy = g * Cfg::GROWS + r;
                    if (y0 + yy < Hi) {
                        float *d = gdst + (long)yy * Wi + 4 * Q;
                        const int D = 4 * Q, lim = min(Cfg::DCOLS, Wi - x0);
                        if (pairs) {   // Wi, x0 and the quad offset are even: a pair never straddles the row limit
                            if (D < lim) red_add_v2(d, sum.x, sum.y);
                            if (D + 2 < lim) red_add_v2(d + 2, sum.z, sum.w);
                        } else {
                            if (D < lim) atomicAdd(d, sum.x);
                            if (D + 1 < lim) atomicAdd(d + 1, sum.y);
                            if (D + 2 < lim) atomicAdd(d + 2, sum.z);
                            if (D + 3 < lim) atomicAdd(d + 3, sum.w);
                        }
                    }
                }
            };

            if (!FOLD) {
#pragma unroll
                for (int r = 0; r < FP; ++r)
                    go[r] = (px < Wo && y0 + r < Ho) ? __ldg(p.gout + ((long)(b * p.C + c) * Ho + y0 + r) * Wo + px) : 0.f;
            }

            constexpr int PRO_CHUNKS = (FP - 2) / Cfg::CH_TAPS + 1;  // chunks touched by the prologue rows
            if (c == 0) {
#pragma unroll
                for (int q = 0; q < PRO_CHUNKS; ++q) mbar_wait(&bars[1 + q], parity);
            }
            static_for<0, FP - 1>([&](auto YY) {
                constexpr int yy = decltype(YY)::value;
                gi_row_v4<KS, 0, yy + 1, (yy > 0), FOLD>(vrow + yy * Cfg::VROW, h, go, ts, st0, stq, wrow(yy - 1), lane);
            });
#pragma unroll
            for (int q = 0; q < Cfg::NCHUNK; ++q) {
                const int lo = max(FP - 1, q * Cfg::CH_TAPS);
                const int hiy = (q == Cfg::NCHUNK - 1) ? KS : min(KS, (q + 1) * Cfg::CH_TAPS);
                if (q >= PRO_CHUNKS && c == 0) mbar_wait(&bars[1 + q], parity);
#pragma unroll 1
                for (int yy = lo; yy < hiy; ++yy) {
                    gi_row_v4<KS, 0, FP, true, FOLD>(vrow + yy * Cfg::VROW, h, go, ts, st0, stq, wrow(yy - 1), lane);
                    if ((yy & 7) == 0) flush((yy >> 3) - 1);  // row yy-1, the last of its group, has just been stored
                }
            }
            static_for<0, FP - 1>([&](auto EI) {
                constexpr int yy = KS + decltype(EI)::value;
                gi_row_v4<KS, decltype(EI)::value + 1, FP, true, FOLD>(vrow + yy * Cfg::VROW, h, go, ts, st0, stq,
                                                                      wrow(yy - 1), lane);
                if ((yy & 7) == 0) flush((yy >> 3) - 1);
            });
            {   // drain the pipeline: the last destination row, then the last group
                constexpr int yy = Cfg::DROWS - 1;
                GiV4Diag g;
                gi4_diag_load(ts, lane, g);
                gi4_diag_store<KS>(g, wrow(yy), lane);
                flush(yy >> 3);
            }
        }
        parity ^= 1;
    }
}

}  // namespace tai
