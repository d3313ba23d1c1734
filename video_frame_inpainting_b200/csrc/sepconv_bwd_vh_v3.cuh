// Gradients of the separable convolution w.r.t. both kernel maps, one pass, for sm_100a
// (persistent, TMA-fed; compile-time ks).
//
//   gV[b,i,y,x] = sum_c gO[b,c,y,x] * sum_j H[b,j,y,x] * I[b,c,y+i,x+j]            (kernel.cu:49-86)
//   gH[b,j,y,x] = sum_c gO[b,c,y,x] * sum_i V[b,i,y,x] * I[b,c,y+i,x+j]            (kernel.cu:88-118)
//
// The reference walks the ks x ks window twice (two kernels, one thread per output element).  Here one
// sweep produces both: a warp owns 8 columns x 4 rows, lane = (cx = lane&7, ch = lane>>3), lane group ch
// owns the horizontal taps j == ch (mod 4).  Per input row yy and output row r (vertical tap i = yy-r):
//     s_r     = sum_{j in group} H_j * I[yy][x+j]      -> gV_i = sum_c gO_c * (s_r summed over the 4 groups)
//     a_r[j] += (V_i * gO_c) * I[yy][x+j]               -> gH_j, complete inside the lane
// The same shared-memory word of I feeds both FMAs.  The four partial s_r are combined by a 3-shuffle
// reduce-scatter after which lane group ch holds the total of output row r = ch and stores it.
// Data movement: as in sepconv_fwd_v3.cuh the H box arrives by TMA and is copied to registers, the V box
// streams in as three tap chunks, the halo is staged by LDG/STS with the replication pad optionally folded in.
// The 4-row boxes are half the size of the forward kernel's, so H and V get a slab EACH (2 x 26 KB + 18 KB of
// halo still fit three CTAs per SM): the V chunks of a tile are requested at its very start (their slab was
// released by the barrier that ended the previous tile) and the H box of the NEXT tile is requested as soon as
// this tile's taps are in registers -- it has the whole sweep to arrive.  Neither TMA latency is on the
// critical path any more (round 1 time-shared one slab: H wait + V wait were serial phases of every tile).
#pragma once

#include "common.cuh"
#include "sepconv_common.cuh"
#include "tma.cuh"

namespace tai {

constexpr int BP = 4;  // output rows per thread

struct BwdParams {
    const float *gout;  // [B,C,Ho,Wo]
    const float *in;    // [B,C,Hi,Wi] (PAD: [B,C,Ho,Wo])
    const float *ver;   // [B,ks,Ho,Wo]
    const float *hor;
    float *gver;        // [B,ks,Ho,Wo] or null
    float *ghor;        // [B,ks,Ho,Wo] or null
    float *gin;         // [B,C,Hi,Wi] or null
    int B, C, Ho, Wo, ks;
    int ntx, nty;
};

template <int KS>
struct VhV3Cfg {
    static constexpr int J = (KS + 3) / 4;
    static constexpr int WX = 4;
    static constexpr int NT = 32 * WX;
    static constexpr int TILE_W = WX * FNX, TILE_H = BP;
    static constexpr int PITCH = TILE_W + 4 * J;
    static constexpr int ROWS = TILE_H + KS - 1;
    static constexpr int NCHUNK = 3;
    static constexpr int CH_TAPS = (KS + NCHUNK - 1) / NCHUNK;
    static constexpr int VROW = TILE_H * TILE_W;
    static constexpr int SLAB_FLOATS = NCHUNK * CH_TAPS * VROW;
    static constexpr int NBAR = 1 + NCHUNK;
    static constexpr int HSLAB_FLOATS = KS * VROW;                 // the H box has a slab of its own (see the kernel)
    static constexpr size_t smem_bytes(int cg) { return (size_t)(SLAB_FLOATS + HSLAB_FLOATS + cg * ROWS * PITCH) * 4 + 8 * NBAR; }
};

struct VhV3Maps {
    CUtensorMap h;  // box {32, KS, 4, 1}, 128-byte swizzle (make_kernel_map_tmap_swz)
    CUtensorMap v;  // box {32, 4, CH_TAPS, 1}
};

// One input row: output rows [RLO, RHI) of this thread are inside the kernel window.
// CG == 1 (FOLD): gO is a per-pixel scalar, so it is multiplied into the H taps once per tile (gV = sum_j (gO H_j) I)
// and into the gH accumulators once at the end (gH_j = gO sum_i V_i I): 8 instructions fewer per row.
template <int KS, int CG, int RLO, int RHI>
__device__ __forceinline__ void vh_row_v3(const float *__restrict__ srow, const float *__restrict__ vrow,
                                          const float (&h)[BP][(KS + 3) / 4], float (&a)[BP][(KS + 3) / 4],
                                          const float (&go)[CG][BP], float (&tsum)[BP])
{
    using Cfg = VhV3Cfg<KS>;
    constexpr int J = Cfg::J;
    constexpr int CSTRIDE = Cfg::ROWS * Cfg::PITCH;
    float v[BP];
#pragma unroll
    for (int r = 0; r < BP; ++r) {
        v[r] = (r >= RLO && r < RHI) ? vrow[r * (Cfg::TILE_W - Cfg::VROW)] : 0.f;  // tap yy-r, output row r
        tsum[r] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < CG; ++c) {
        float iv[J];
#pragma unroll
        for (int jj = 0; jj < J; ++jj) iv[jj] = srow[c * CSTRIDE + 4 * jj];
        // tap-outer order: the row sums s[r] are serial FMA chains; interleaving the rows (and the independent
        // a[r][jj] updates) keeps dependent FMAs 2*(RHI-RLO) issue slots apart
        float w[BP], s[BP];
#pragma unroll
        for (int r = RLO; r < RHI; ++r) {
            w[r] = CG == 1 ? v[r] : v[r] * go[c][r];
            s[r] = h[r][0] * iv[0];
            a[r][0] = fmaf(w[r], iv[0], a[r][0]);
        }
#pragma unroll
        for (int jj = 1; jj < J; ++jj)
#pragma unroll
            for (int r = RLO; r < RHI; ++r) {
                s[r] = fmaf(h[r][jj], iv[jj], s[r]);
                a[r][jj] = fmaf(w[r], iv[jj], a[r][jj]);
            }
#pragma unroll
        for (int r = RLO; r < RHI; ++r) tsum[r] = CG == 1 ? s[r] : fmaf(go[c][r], s[r], tsum[r]);
    }
}

template <int KS, int CG, bool PAD>
// small windows keep few taps in registers: six CTAs per SM for ks <= 16 (80 registers; ks = 13 backward + 10 %), four for
// ks <= 28 (ks = 25: + 4 %) put more boxes in flight where the op is HBM-bound
__global__ void __launch_bounds__(128, (KS <= 16 ? 6 : KS <= 28 ? 4 : 3))
sepconv_bwd_vh_v3_kernel(const __grid_constant__ VhV3Maps maps, const BwdParams p)
{
    static_assert(BP == 4, "the reduce-scatter assumes 4 output rows == 4 tap groups");
    using Cfg = VhV3Cfg<KS>;
    constexpr int J = Cfg::J, PITCH = Cfg::PITCH, ROWS = Cfg::ROWS, TILE_W = Cfg::TILE_W, TILE_H = Cfg::TILE_H;
    constexpr int CSTRIDE = ROWS * PITCH;
    extern __shared__ __align__(1024) float smem[];
    float *hslab = smem;                                 // H box of the current tile, then of the next one (swizzled: 1024-byte aligned)
    float *slab = smem + Cfg::HSLAB_FLOATS;              // V chunks of the current tile
    float *is = slab + Cfg::SLAB_FLOATS;
    uint64_t *bars = reinterpret_cast<uint64_t *>(is + CG * CSTRIDE);

    const int Ho = p.Ho, Wo = p.Wo;
    const int Hi = Ho + KS - 1, Wi = Wo + KS - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cx = lane & 7, ch = lane >> 3;
    const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
    const long plane = (long)Ho * Wo;
    const int ntiles = p.B * p.nty * p.ntx;

    // Tiles are numbered with the ROW index fastest and every CTA walks a contiguous range, so the next tile is
    // usually the one directly below: it shares ks-1 of its ks+3 halo rows with the current one, and only the 4
    // new rows are staged (cp.async) into the ring slots the sweep no longer needs -- the halo staging of 54 rows
    // per 4 output rows was the longest phase between two sweeps.
    // tile index -> (x0, y0, b); tiles that would stick out are shifted back inside
    auto tile_origin = [&](int tile, int &x0, int &y0, int &b) {
        const int ty = tile % p.nty;
        tile /= p.nty;
        x0 = max(0, min((tile % p.ntx) * TILE_W, Wo - TILE_W));
        y0 = min(ty * TILE_H, Ho - TILE_H);  // host guarantees Ho >= TILE_H
        b = tile / p.ntx;
    };
    const int t_lo = (int)((long)ntiles * blockIdx.x / gridDim.x), t_hi = (int)((long)ntiles * (blockIdx.x + 1) / gridDim.x);

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < Cfg::NBAR; ++i) mbar_init(&bars[i], 1);
        mbar_fence_init();
        if ((__cvta_generic_to_shared(hslab) & 1023) != 0) __trap();  // the swizzle pattern is tied to 1024-byte blocks
        if (t_lo < t_hi) {   // H box of this CTA's first tile
            int fx, fy, fb;
            tile_origin(t_lo, fx, fy, fb);
            mbar_expect_tx(&bars[0], KS * Cfg::VROW * 4);
            tma_load_4d(hslab, &maps.h, &bars[0], fx, 0, fy, fb);
        }
    }
    __syncthreads();
    uint32_t parity = 0;
    int swz[8];
    swz_table(warp * FNX + cx, ch, swz);
    int rbase = 0;                          // ring row that holds halo row 0 of the current tile
    int px0 = -1, py0 = -1, pb = -1;        // origin of the previous tile

    for (int tile = t_lo; tile < t_hi; ++tile) {
        int x0, y0, b;
        tile_origin(tile, x0, y0, b);
        const bool walk = (b == pb && x0 == px0 && y0 == py0 + TILE_H);
        px0 = x0; py0 = y0; pb = b;
        const int px_raw = x0 + warp * FNX + cx;
        const bool px_ok = px_raw < Wo;
        const int px = px_ok ? px_raw : Wo - 1;

        if (threadIdx.x == 0) {   // V chunks of this tile: their slab was released by the barrier that ended the last tile
            fence_proxy_async();
#pragma unroll
            for (int q = 0; q < Cfg::NCHUNK; ++q) {
                mbar_expect_tx(&bars[1 + q], Cfg::CH_TAPS * Cfg::VROW * 4);
                tma_load_4d(slab + q * Cfg::CH_TAPS * Cfg::VROW, &maps.v, &bars[1 + q], x0, y0, q * Cfg::CH_TAPS, b);
            }
        }
        // ---- halo of all CG (== C) channels: everything for a new column, the 4 new rows when walking down ----
        {
            constexpr int NK = (PITCH + 31) / 32;
            int coff[NK];
            bool cok[NK];
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                const int rx = lane + 32 * k, gx = x0 + rx;
                if (PAD) {
                    cok[k] = rx < TILE_W + KS - 1;
                    coff[k] = clampi(gx - KS / 2, 0, Wo - 1);
                } else {
                    cok[k] = rx < TILE_W + KS - 1 && gx < Wi;
                    coff[k] = cok[k] ? gx : 0;
                }
            }
            if (walk) {
                rbase += TILE_H;
                if (rbase >= ROWS) rbase -= ROWS;
                // halo rows ROWS-4..ROWS-1 of this tile -> the ring slots of the previous tile's rows 0..3 (asynchronous:
                // they are first read at sweep row ROWS-4; waited for at the start of the last V chunk)
                for (int c = 0; c < CG; ++c) {
                    const float *src = PAD ? p.in + ((long)(b * CG + c)) * plane : p.in + ((long)(b * CG + c)) * Hi * Wi;
                    for (int ry = ROWS - TILE_H + warp; ry < ROWS; ry += Cfg::NT / 32) {
                        const int gy = y0 + ry;
                        const float *grow = PAD ? src + (long)clampi(gy - KS / 2, 0, Ho - 1) * Wo : src + (long)gy * Wi;
                        int rr = rbase + ry;
                        if (rr >= ROWS) rr -= ROWS;
                        float *drow = is + c * CSTRIDE + rr * PITCH + lane;
#pragma unroll
                        for (int k = 0; k < NK; ++k)
                            if (lane + 32 * k < PITCH) cp_async_f32(drow + 32 * k, grow + coff[k], cok[k]);
                    }
                }
                cp_async_commit();
            } else {
                rbase = 0;
                constexpr int RB = (ROWS + Cfg::NT / 32 - 1) / (Cfg::NT / 32);  // all rows of a warp in one batch: one memory round trip
                for (int c = 0; c < CG; ++c) {
                    const float *src = PAD ? p.in + ((long)(b * CG + c)) * plane : p.in + ((long)(b * CG + c)) * Hi * Wi;
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
        }
        float go[CG][BP];
#pragma unroll
        for (int c = 0; c < CG; ++c)
#pragma unroll
            for (int r = 0; r < BP; ++r) go[c][r] = __ldg(p.gout + ((long)(b * CG + c) * Ho + y0 + r) * Wo + px);

        // ---- H taps: slab -> registers (the box was requested during the previous tile) ----
        mbar_wait(&bars[0], parity);
        float h[BP][J], a[BP][J];
        {
#pragma unroll
            for (int jj = 0; jj < J; ++jj)
#pragma unroll
                for (int r = 0; r < BP; ++r) {
                    h[r][jj] = (ch + 4 * jj < KS) ? hslab[(r * KS + 4 * jj) * 32 + swz[(r * KS + 4 * jj) & 7]] : 0.f;
                    if (CG == 1) h[r][jj] *= go[0][r];
                    a[r][jj] = 0.f;
                }
        }
        __syncthreads();  // H is in registers everywhere; the halo is complete
        if (threadIdx.x == 0 && tile + 1 < t_hi) {
            // the H box of the next tile into the slab that has just been read out, its V chunks into L2
            int nx0, ny0, nb;
            tile_origin(tile + 1, nx0, ny0, nb);
            fence_proxy_async();
            mbar_expect_tx(&bars[0], KS * Cfg::VROW * 4);
            tma_load_4d(hslab, &maps.h, &bars[0], nx0, 0, ny0, nb);
#pragma unroll
            for (int q = 0; q < Cfg::NCHUNK; ++q) tma_prefetch_l2_4d(&maps.v, nx0, ny0, q * Cfg::CH_TAPS, nb);
        }

        const float *sbase = is + warp * FNX + cx + ch;
        const float *vrow = slab + warp * FNX + cx;
        float *gv = p.gver ? p.gver + ((long)b * KS * Ho + y0 + ch) * Wo + px : nullptr;  // + tap * plane
        auto ring_row = [&](int yy) {   // halo row yy of this tile inside the ring
            int rr = rbase + yy;
            if (rr >= ROWS) rr -= ROWS;
            return sbase + rr * PITCH;
        };

        const bool gv_ok = gv != nullptr && px_ok;
        // After one input row: combine the four tap groups (lane group ch ends up with the total of output row ch) ...
        auto reduce_rows = [&](float (&tsum)[BP]) {
            const float keep0 = hi16 ? tsum[2] : tsum[0];
            const float keep1 = hi16 ? tsum[3] : tsum[1];
            const float send0 = hi16 ? tsum[0] : tsum[2];
            const float send1 = hi16 ? tsum[1] : tsum[3];
            const float u0 = keep0 + __shfl_xor_sync(0xffffffffu, send0, 16);
            const float u1 = keep1 + __shfl_xor_sync(0xffffffffu, send1, 16);
            return (hi8 ? u1 : u0) + __shfl_xor_sync(0xffffffffu, hi8 ? u0 : u1, 8);
        };
        // ... and store gV[tap yy-ch][row ch] (prologue / epilogue rows: the tap may fall outside [0, ks))
        auto finish_row = [&](int yy, float (&tsum)[BP]) {
            const float tot = reduce_rows(tsum);
            const int i = yy - ch;
            if (gv_ok && i >= 0 && i < KS) gv[(long)i * plane] = tot;
        };

        constexpr int PRO_CHUNKS = (BP - 2) / Cfg::CH_TAPS + 1;  // chunks touched by the prologue rows
#pragma unroll
        for (int q = 0; q < PRO_CHUNKS; ++q) mbar_wait(&bars[1 + q], parity);
        static_for<0, BP - 1>([&](auto YY) {
            constexpr int yy = decltype(YY)::value;
            float tsum[BP];
            vh_row_v3<KS, CG, 0, yy + 1>(ring_row(yy), vrow + yy * Cfg::VROW, h, a, go, tsum);
            finish_row(yy, tsum);
        });
#pragma unroll
        for (int q = 0; q < Cfg::NCHUNK; ++q) {
            const int lo = max(BP - 1, q * Cfg::CH_TAPS);
            const int hi = (q == Cfg::NCHUNK - 1) ? KS : min(KS, (q + 1) * Cfg::CH_TAPS);
            if (q >= PRO_CHUNKS) mbar_wait(&bars[1 + q], parity);
            if (q == Cfg::NCHUNK - 1 && walk) {   // the 4 new halo rows are first read at row ROWS-4 >= lo
                cp_async_wait_all();
                __syncthreads();
            }
            const float *srow = ring_row(lo);
            const float *wrap = sbase + ROWS * PITCH;
            // steady rows: every lane's tap yy - ch is inside [0, ks), the store pointer just walks one tap plane per row
            float *gvp = gv_ok ? gv + (long)(lo - ch) * plane : nullptr;
            if constexpr (CG == 1) {
                // two rows per trip (half the loop overhead); the V row pointer walks with the halo row pointer
                const float *vr = vrow + lo * Cfg::VROW;
                auto steady_row = [&]() {
                    float tsum[BP];
                    vh_row_v3<KS, CG, 0, BP>(srow, vr, h, a, go, tsum);
                    const float tot = reduce_rows(tsum);
                    if (gv_ok) {
                        *gvp = tot;
                        gvp += plane;
                    }
                    vr += Cfg::VROW;
                    srow += PITCH;
                    if (srow >= wrap) srow -= ROWS * PITCH;
                };
                int yy = lo;
#pragma unroll 1
                for (; yy + 1 < hi; yy += 2) {
                    steady_row();
                    steady_row();
                }
                if (yy < hi) steady_row();
            } else {
                // C = 3: the body is 2.6x longer; the unrolled form measured 2-5 % slower (0.556 -> 0.57-0.58 ms at the UCF shape)
#pragma unroll 1
                for (int yy = lo; yy < hi; ++yy) {
                    float tsum[BP];
                    vh_row_v3<KS, CG, 0, BP>(srow, vrow + yy * Cfg::VROW, h, a, go, tsum);
                    const float tot = reduce_rows(tsum);
                    if (gv_ok) {
                        *gvp = tot;
                        gvp += plane;
                    }
                    srow += PITCH;
                    if (srow >= wrap) srow -= ROWS * PITCH;
                }
            }
        }
        static_for<0, BP - 1>([&](auto E) {
            constexpr int yy = KS + decltype(E)::value;
            float tsum[BP];
            vh_row_v3<KS, CG, decltype(E)::value + 1, BP>(ring_row(yy), vrow + yy * Cfg::VROW, h, a, go, tsum);
            finish_row(yy, tsum);
        });
        parity ^= 1;

        if (p.ghor && px_ok) {
            if constexpr (CG == 1) {
                float *gh = p.ghor + (((long)b * KS + ch) * Ho + y0) * Wo + px;   // tap ch; the pointer walks 4 tap planes per slot
                const long step = 4 * plane;
#pragma unroll
                for (int jj = 0; jj < J; ++jj) {
                    if (ch + 4 * jj < KS) {
#pragma unroll
                        for (int r = 0; r < BP; ++r) gh[r * Wo] = a[r][jj] * go[0][r];
                    }
                    gh += step;
                }
            } else {
                float *gh = p.ghor + ((long)b * KS * Ho + y0) * Wo + px;
#pragma unroll
                for (int jj = 0; jj < J; ++jj) {
                    const int j = ch + 4 * jj;
                    if (j < KS) {
#pragma unroll
                        for (int r = 0; r < BP; ++r) gh[(long)j * plane + (long)r * Wo] = a[r][jj];
                    }
                }
            }
        }
        __syncthreads();  // slab and halo are free for the next tile
    }
}

}  // namespace tai
