// Forward separable convolution for sm_100a, fifth layout: persistent CTAs, a TMA CHUNK RING for the two
// kernel maps and a ROW RING for the input halo (compile-time ks).
//
//   O[b,c,y,x] = sum_i V[b,i,y,x] * ( sum_j H[b,j,y,x] * I[b,c,y+i,x+j] )          (kernel.cu:19-47)
//
// Work decomposition (unchanged from v3): CTA = 4 warps, tile = 8 rows x 32 columns, a warp owns
// 8 columns x 8 rows; lane = (cx = lane&7, ch = lane>>3); lane group ch keeps the horizontal taps
// j == ch (mod 4) of its 8 pixels in registers and sweeps the ks+7 input rows of the tile.
//
// What is new
//   * CHUNK RING.  A kernel-map box [ks taps][8 rows][32 cols] is cut into NCH chunks of 8 taps (8 KB); the
//     ring has NCH slots and slot k holds, in turn, H chunk k and V chunk k of every tile.  Each slot has
//     a FULL mbarrier (TMA transaction bytes) and a release counter.  There is no producer thread and no
//     polling: a warp that has finished with a slot bumps the counter, and the warp that arrives LAST
//     issues the slot's next TMA load itself (H chunk k -> V chunk k of the same tile -> H chunk k of the
//     next tile).  Every load is therefore issued at the first moment its slot is free: the H chunks of
//     tile n+1 stream in while tile n is still being filtered, the V chunks follow as soon as the H taps
//     are in registers, and nothing waits for a whole box.  (Measured alternative: one thread polling
//     EMPTY mbarriers at chunk boundaries cost its warp ~800 cycles per visit and stalled the CTA.)
//   * ROW RING.  Tiles are numbered with the row index fastest and every CTA takes a contiguous range
//     (balanced per SM first, then over the CTAs of the SM), so the next tile is usually the one
//     directly below: it shares ks-1 of its ks+7 halo rows with the current one and only the 8 new rows
//     are staged (cp.async), into the ring rows the previous tile no longer needs.
//   * DUAL (the bi-TAI call site) makes one pass over the range per prediction stream; pass 0 writes Dot1
//     and a*Dot1, pass 1 writes Dot2 and adds b*Dot2 (tai.py:105 / twi.py:105): each thread re-reads
//     exactly the element it wrote itself.  With PAD the replication pad of tai.py:170-171 is folded
//     into the halo staging (clamped source coordinates).
#pragma once

#include "common.cuh"
#include "sepconv_common.cuh"
#include "tma.cuh"

namespace tai {

#ifdef TAI_LAB_TIMING
__device__ unsigned long long g_lab_phase[16];
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
struct FwdV5Cfg {
    static constexpr int CT = 8;                       // taps per chunk
    static constexpr int NCH = (KS + CT - 1) / CT;     // chunks per box == ring slots
    static constexpr int J = (KS + 3) / 4;             // taps per lane
    static constexpr int WX = 4, NT = 32 * WX;
    static constexpr int TILE_W = WX * FNX, TILE_H = FP;
    static constexpr int PITCH = TILE_W + KS - 1;      // halo row, no padding (a row is read by one LDS at a time)
    static constexpr int ROWS = TILE_H + KS - 1;
    static constexpr int VROW = TILE_H * TILE_W;       // floats per tap in a chunk
    static constexpr int CHUNK_FLOATS = CT * VROW;
    static constexpr int SLAB_FLOATS = NCH * CHUNK_FLOATS;
    static constexpr int TAIL = 4;                     // lanes of tap group 3 read up to 3 floats past a row
    static constexpr size_t smem_bytes(int cg)
    {
        return (size_t)(SLAB_FLOATS + cg * ROWS * PITCH + TAIL) * 4 + 16 * NCH + 16 * 4 * 4;  // + FULL barriers, counters, tile scratch
    }
    static_assert(NCH >= 2, "the kernel needs at least two chunks");
    static_assert((ROWS * PITCH) % 4 == 0, "barrier alignment");
};

struct FwdV5Maps {
    CUtensorMap h[2];   // box {32, 8, CT, 1}: the chunk loads
    CUtensorMap v[2];
    CUtensorMap hb[2];  // box {32, 8, KS, 1}: whole-box L2 prefetch of the tile after next
    CUtensorMap vb[2];
};

// One chunk load (kept out of line: it is instantiated at every release site).
__device__ __noinline__ void fwd_v5_issue(float *dst, const CUtensorMap *tm, uint64_t *bar, uint32_t bytes, int x0,
                                          int y0, int tap0, int b)
{
    mbar_expect_tx(bar, bytes);
    tma_load_4d(dst, tm, bar, x0, y0, tap0, b);
}

// One input row of the sweep for output rows [RLO, RHI) of this thread.
template <int KS, int CG, int RLO, int RHI>
__device__ __forceinline__ void fwd_row_v5(const float *__restrict__ srow, const float *__restrict__ vrow,
                                           const float (&h)[FP][(KS + 3) / 4], float (&acc)[CG][FP])
{
    using Cfg = FwdV5Cfg<KS>;
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
        // tap-outer order: the RHI-RLO row sums are independent FMA chains that interleave (a row-outer order
        // makes ptxas emit one serial chain after the other: 4-cycle dependent issue with 3 warps per scheduler)
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
__global__ void __launch_bounds__(128, (CG == 1 ? 3 : 2))
sepconv_fwd_v5_kernel(const __grid_constant__ FwdV5Maps maps, const FwdParams p, const int cps)
{
    using Cfg = FwdV5Cfg<KS>;
    constexpr int J = Cfg::J, PITCH = Cfg::PITCH, ROWS = Cfg::ROWS, TILE_W = Cfg::TILE_W, TILE_H = Cfg::TILE_H;
    constexpr int NCH = Cfg::NCH, CT = Cfg::CT, VROW = Cfg::VROW;
    constexpr int CSTRIDE = ROWS * PITCH;
    constexpr int NS = DUAL ? 2 : 1;
    constexpr int NK = (PITCH + 31) / 32;
    constexpr uint32_t CHUNK_BYTES = Cfg::CHUNK_FLOATS * 4;
    extern __shared__ __align__(1024) float smem[];
    float *slab = smem;                       // [NCH][CT][TILE_H][TILE_W], written by TMA only
    float *is = smem + Cfg::SLAB_FLOATS;      // [CG][ROWS (ring)][PITCH] + TAIL
    uint64_t *full = reinterpret_cast<uint64_t *>(is + CG * CSTRIDE + Cfg::TAIL);
    unsigned *cnt = reinterpret_cast<unsigned *>(full + NCH);  // releases of slot k so far
    // per-warp tile scratch: the coordinates are needed only by the rare refill path and by the stores after
    // the sweep; parking them in shared memory keeps ~12 registers free for the sweep's FMA chains
    int *meta = reinterpret_cast<int *>(cnt + 2 * NCH) + 16 * (threadIdx.x >> 5);

    const int Ho = p.Ho, Wo = p.Wo;
    const int Hi = Ho + KS - 1, Wi = Wo + KS - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cx = lane & 7, ch = lane >> 3;
    const long plane = (long)Ho * Wo;

    // ---- this CTA's contiguous tile range: balanced per SM, then over the SM's CTAs -----------------
    // tile id = (b * ntx + tx) * nty + ty
    const int ntiles = p.B * p.nty * p.ntx;
    const int nsm = gridDim.x / cps;
    const int sm = blockIdx.x % nsm, sl = blockIdx.x / nsm;
    const int sm_lo = (int)((long)ntiles * sm / nsm), sm_hi = (int)((long)ntiles * (sm + 1) / nsm);
    const int t_lo = sm_lo + (sm_hi - sm_lo) * sl / cps;
    const int t_hi = sm_lo + (sm_hi - sm_lo) * (sl + 1) / cps;
    if (t_lo >= t_hi) return;
    const int npass = (p.C / CG) * NS;        // pass = (channel group, stream)

    // ---- tile walker: coordinates of consecutive tile ids without divisions in the loop ----------------
    struct Walk {
        int ty, tx, b;
    };
    auto walk_next = [&](Walk &w) {
        if (++w.ty == p.nty) {
            w.ty = 0;
            if (++w.tx == p.ntx) {
                w.tx = 0;
                ++w.b;
            }
        }
    };
    Walk w_lo;
    w_lo.ty = t_lo % p.nty;
    w_lo.tx = (t_lo / p.nty) % p.ntx;
    w_lo.b = (t_lo / p.nty) / p.ntx;
    auto tile_x0 = [&](const Walk &w) { return max(0, min(w.tx * TILE_W, Wo - TILE_W)); };
    auto tile_y0 = [&](const Walk &w) { return min(w.ty * TILE_H, Ho - TILE_H); };  // host guarantees Ho >= TILE_H

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            mbar_init(&full[i], 1);
            cnt[i] = 0u;
        }
        mbar_fence_init();
#pragma unroll
        for (int i = 0; i < Cfg::TAIL; ++i) is[CG * CSTRIDE + i] = 0.f;
        // the H chunks of the first item; everything after that is issued by the last releaser of a slot
#pragma unroll
        for (int k = 0; k < NCH; ++k)
            fwd_v5_issue(slab + k * Cfg::CHUNK_FLOATS, &maps.h[0], &full[k], CHUNK_BYTES, tile_x0(w_lo), tile_y0(w_lo),
                         k * CT, w_lo.b);
    }
    __syncthreads();
#ifdef TAI_LAB_TIMING
    long long lab_t_ = clock64();
#endif

    for (int pass = 0; pass < npass; ++pass) {
        const int s = pass % NS;
        const int c0 = (pass / NS) * CG;
        const float *__restrict__ in = s ? p.in[1] : p.in[0];
        float *__restrict__ out_s = s ? p.out[1] : p.out[0];
        const CUtensorMap *tm_v = s ? &maps.v[1] : &maps.v[0];
        int rbase = 0;                       // ring row that holds halo row 0 of the current tile
        Walk cw = w_lo;

#pragma unroll 1
        for (int tile = t_lo; tile < t_hi; ++tile) {
            int x0, y0, b;
            bool walk;
            {
                const int ty = cw.ty, tx = cw.tx;
                b = cw.b;
                // Tiles that would stick out are shifted back inside (the overlap recomputes identical values).
                x0 = tile_x0(cw);
                y0 = tile_y0(cw);
                // the tile directly below the previous one of this pass shares ks-1 halo rows with it
                walk = tile > t_lo && ty > 0 && ty * TILE_H + TILE_H <= Ho;
                // the item after this one (its H chunks are issued as this item's V slots are released)
                const bool last_of_pass = tile + 1 == t_hi;
                const bool has_next = !last_of_pass || pass + 1 < npass;
                walk_next(cw);
                if (last_of_pass) cw = w_lo;
                if (lane == 0) {
                    meta[0] = x0;
                    meta[1] = y0;
                    meta[2] = b;
                    meta[3] = tile_x0(cw);
                    meta[4] = tile_y0(cw);
                    meta[5] = cw.b;
                    meta[6] = has_next ? ((NS == 2 && (last_of_pass ? !s : s)) ? 2 : 1) : 0;  // 0: none, 1: stream 0, 2: stream 1
                    meta[7] = tx * TILE_W;       // first column of the tile's own grid cell
                    meta[8] = ty * TILE_H - y0;  // first output row the tile owns (0 unless shifted)
                }
                __syncwarp();
            }

            // Slot hand-over, split in two so that the warp never waits for the atomic's round trip:
            //   arrive(k)  after the warp's last read of slot k: lane 0 bumps the slot's release counter;
            //   settle()   a few hundred cycles later: if this warp was the LAST of the four to arrive, it refills
            //              the slot (H chunk k -> V chunk k of this tile -> H chunk k of the next tile).
            unsigned pend_old = 0;
            int pend_k = -1;
            bool pend_v = false;
            auto settle = [&]() {
                if (lane == 0 && pend_k >= 0 && (pend_old & (Cfg::WX - 1)) == Cfg::WX - 1) {
#ifdef TAI_LAB_TIMING
                    const long long t0_ = clock64();
#endif
                    const volatile int *m = meta;
                    if (pend_v)
                        fwd_v5_issue(slab + pend_k * Cfg::CHUNK_FLOATS, tm_v, &full[pend_k], CHUNK_BYTES, m[0], m[1],
                                     pend_k * CT, m[2]);
                    else if (m[6])
                        fwd_v5_issue(slab + pend_k * Cfg::CHUNK_FLOATS, m[6] == 2 ? &maps.h[1] : &maps.h[0],
                                     &full[pend_k], CHUNK_BYTES, m[3], m[4], pend_k * CT, m[5]);
#ifdef TAI_LAB_TIMING
                    if (threadIdx.x == 0) {
                        atomicAdd(&g_lab_phase[5], (unsigned long long)(clock64() - t0_));
                        atomicAdd(&g_lab_phase[6], 1ull);
                    }
#endif
                }
                pend_k = -1;
            };
            auto arrive = [&](int k, bool next_is_v) {
                settle();
                __syncwarp();
                if (lane == 0) pend_old = atomicAdd(&cnt[k], 1u);
                pend_k = k;
                pend_v = next_is_v;
            };

            // ---- halo rows -> ring (all of them for a new column, the 8 new ones when walking down) ----
            {
                int first_row = 0;
                if (walk) {
                    rbase += TILE_H;
                    if (rbase >= ROWS) rbase -= ROWS;
                    first_row = ROWS - TILE_H;
                } else {
                    rbase = 0;
                    __syncthreads();  // every warp has left the previous tile: the whole ring may be overwritten
                }
                int coff[NK];
                bool cok[NK];
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    const int rx = lane + 32 * k, gx = x0 + rx;
                    if (PAD) {
                        cok[k] = rx < PITCH;
                        coff[k] = clampi(gx - KS / 2, 0, Wo - 1);
                    } else {
                        cok[k] = rx < PITCH && gx < Wi;
                        coff[k] = cok[k] ? gx : 0;
                    }
                }
                for (int c = 0; c < CG; ++c) {
                    const float *src = PAD ? in + ((long)(b * p.C + c0 + c)) * plane
                                           : in + ((long)(b * p.C + c0 + c)) * Hi * Wi;
                    for (int ry = first_row + warp; ry < ROWS; ry += Cfg::WX) {
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
                if (!walk) {
                    cp_async_wait_all();
                    __syncthreads();
                }
            }
            LAB_T(0);

            // DUAL, second stream: fetch a*Dot1 (written by this very thread in pass 0) now, use it after the sweep
            float part[CG][FP / 4];
            if (DUAL && s == 1) {
                const int pxr = x0 + warp * FNX + cx;
                if (pxr < Wo) {
#pragma unroll
                    for (int c = 0; c < CG; ++c)
#pragma unroll
                        for (int q = 0; q < FP / 4; ++q)
                            part[c][q] = __ldcg(p.blend + ((long)(b * p.C + c0 + c) * Ho + y0 + ch + 4 * q) * Wo + pxr);
                }
            }

            // ---- this lane's horizontal taps: H chunks -> registers, chunk by chunk ----
            float h[FP][J];
            static_for<0, NCH>([&](auto K) {
                constexpr int k = decltype(K)::value;
                mbar_wait(&full[k], 0);
                const float *hs = slab + k * Cfg::CHUNK_FLOATS + ch * VROW + warp * FNX + cx;
#pragma unroll
                for (int r = 0; r < FP; ++r) {
                    if (2 * k < J) h[r][2 * k] = hs[r * TILE_W];
                    if (2 * k + 1 < J) h[r][2 * k + 1] = hs[4 * VROW + r * TILE_W];
                }
                arrive(k, true);
            });
            settle();  // a pending V refill must not wait for the sweep's first hand-over (the sweep may need it first)
            LAB_T(1);

            float acc[CG][FP];
#pragma unroll
            for (int c = 0; c < CG; ++c)
#pragma unroll
                for (int r = 0; r < FP; ++r) acc[c][r] = 0.f;

            const float *sbase = is + warp * FNX + cx + ch;
            const float *vrow = slab + warp * FNX + cx;  // tap 0, row 0
            auto ring_row = [&](int yy) {                // halo row yy of this tile inside the ring
                int rr = rbase + yy;
                if (rr >= ROWS) rr -= ROWS;
                return sbase + rr * PITCH;
            };

            // ---- sweep.  Row yy touches the vertical taps yy-7..yy: chunk q is first needed at row 8q ----
            mbar_wait(&full[0], 1);
            LAB_T(2);
            static_for<0, FP - 1>([&](auto YY) {        // prologue: output rows 0..yy are inside the window
                constexpr int yy = decltype(YY)::value;
                fwd_row_v5<KS, CG, 0, yy + 1>(ring_row(yy), vrow + yy * VROW, h, acc);
            });
            fwd_row_v5<KS, CG, 0, FP>(ring_row(FP - 1), vrow + (FP - 1) * VROW, h, acc);
            static_for<1, NCH>([&](auto Q) {
                constexpr int q = decltype(Q)::value;
                LAB_T(12);
                mbar_wait(&full[q], 1);
                LAB_T(13);
                if (q == NCH - 1) {      // the 8 new halo rows are first read at row ks-1, inside this chunk
                    cp_async_wait_all();
                    __syncthreads();
                }
                LAB_T(14);
                constexpr int lo = CT * q, hi = (CT * q + CT < KS) ? CT * q + CT : KS;
                {
                    int rr = rbase + lo;
                    if (rr >= ROWS) rr -= ROWS;
                    const float *srow = sbase + rr * PITCH;
                    const float *wrap = sbase + ROWS * PITCH;
#pragma unroll 1
                    for (int yy = lo; yy < hi; ++yy) {
                        fwd_row_v5<KS, CG, 0, FP>(srow, vrow + yy * VROW, h, acc);
                        srow += PITCH;
                        if (srow >= wrap) srow -= ROWS * PITCH;
                    }
                }
                // chunk q-1 holds taps <= 8q-1; the rows still to come need taps >= hi-7
                LAB_T(12);
                if (hi - (FP - 1) > CT * q - 1) arrive(q - 1, false);
                LAB_T(15);
            });
            static_for<0, FP - 1>([&](auto E) {          // epilogue: input rows ks..ks+6
                constexpr int yy = KS + decltype(E)::value;
                fwd_row_v5<KS, CG, decltype(E)::value + 1, FP>(ring_row(yy), vrow + yy * VROW, h, acc);
            });
            // release the chunks the loop above could not release yet
            static_for<1, NCH + 1>([&](auto Q) {
                constexpr int q = decltype(Q)::value;
                constexpr int hi = (q < NCH) ? ((CT * q + CT < KS) ? CT * q + CT : KS) : KS + FP;
                constexpr bool released = (q < NCH) && (hi - (FP - 1) > CT * q - 1);
                if (!released) arrive(q - 1, false);
            });
            settle();
            LAB_T(3);

            // ---- sum the four tap groups; lane group ch stores rows r == ch (mod 4) ----
            const volatile int *m = meta;
            x0 = m[0];
            y0 = m[1];
            b = m[2];
            const int px_raw = x0 + warp * FNX + cx;
            // a shifted tile stores only the pixels of its own grid cell (pass 1 of DUAL accumulates, so a
            // pixel must be stored by exactly one tile)
            const bool px_ok = px_raw < Wo && px_raw >= m[7];
            const int r_own = m[8];
#pragma unroll
            for (int c = 0; c < CG; ++c)
#pragma unroll
                for (int r = 0; r < FP; ++r) {
                    float a = acc[c][r];
                    a += __shfl_xor_sync(0xffffffffu, a, 8);
                    a += __shfl_xor_sync(0xffffffffu, a, 16);
                    if ((r & 3) == ch && px_ok && r >= r_own) {
                        const long o = ((long)(b * p.C + c0 + c) * Ho + y0 + r) * Wo + px_raw;
                        if (DUAL) {
                            if (out_s) out_s[o] = a;
                            if (s == 0)
                                p.blend[o] = p.a * a;
                            else
                                p.blend[o] = fmaf(p.b, a, part[c][r >> 2]);
                        } else {
                            out_s[o] = a;
                        }
                    }
                }
            LAB_T(4);
        }
        __syncthreads();  // the next pass restages the whole ring
    }
}

}  // namespace tai
