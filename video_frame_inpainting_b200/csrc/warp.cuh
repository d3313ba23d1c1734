// Bilinear backward warp helpers shared by elementwise.cu (FlowWarper) and slomo.cu (fused SloMo stages).
#pragma once

#include "common.cuh"

namespace tai {

static inline unsigned stream_grid(long work_items, int block)
{
    long g = (work_items + block - 1) / block;
    const long cap = (long)sm_count() * 8;  // 8 resident 256-thread CTAs per SM, grid-stride beyond
    if (g > cap) {
        // every thread makes the same number of grid-stride trips (a grid of exactly `cap` CTAs leaves a ragged
        // second trip: 2048 CTAs of work on 1184 slots ran as 1 + 0.73 waves)
        const long trips = (g + cap - 1) / cap;
        g = (g + trips - 1) / trips;
    }
    if (g < 1) g = 1;
    return (unsigned)g;
}

// Grid for a grid-stride kernel whose loop handles `unroll` items per trip: every thread makes the same number
// of trips (a multiple of `unroll`), and the grid fits the kernel's real residency (occupancy query, cached by
// the caller) so that it runs as ONE wave.  Small CTAs keep the per-SM CTA count even (1024 CTAs on 148 SMs:
// 7 vs 6.9 average; 512 larger CTAs: 4 vs 3.5).
template <typename K>
static inline unsigned stream_grid_occ(K kernel, long work_items, int block, int unroll)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    const long cap = (long)sm_count() * per_sm;
    long g = (work_items + block - 1) / block;
    if (g > cap) {
        long trips = (g + cap - 1) / cap;
        trips = (trips + unroll - 1) / unroll * unroll;
        g = (g + trips - 1) / trips;
    }
    if (g < 1) g = 1;
    return (unsigned)g;
}

// ------------------------------------------------------------------------------------------------
// Bilinear backward warp (slomo.py:265-286 + torch-0.3.1 grid_sample: bilinear, zero padding).
// The coordinate chain is evaluated with one IEEE rounding per reference operation (no FMA
// contraction) so that floor() -- the integer part of the op -- matches the FP32 reference exactly:
//   X = x + u;  g = 2*(X/W - 0.5);  ix = ((g + 1)/2)*(W-1)
struct WarpCoord {
    int x0, y0;
    float ix, iy;
};

// Correctly rounded a / b from the correctly rounded reciprocal y = RN(1 / b) (formed once on the host):
// q = RN(a * y), r = a - b * q (exact in an FMA), result = RN(q + r * y)  [Markstein].  b is the image width or
// height -- a small integer --, a is a pixel coordinate: the result equals IEEE division bit for bit (12 M
// random and near-integer cases over 24 sizes checked against FP32 division: no mismatch), in 3 instructions
// instead of the ~10 + slow-path call of __fdiv_rn.  The four divisions of the coordinate chain made the warp
// kernels instruction-bound (~1000 SASS instructions per pixel at C = 3).
__device__ __forceinline__ float div_by_size(float a, float b, float y)
{
    const float q = __fmul_rn(a, y);
    const float r = __fmaf_rn(-b, q, a);
    return __fmaf_rn(r, y, q);
}

__device__ __forceinline__ float warp_axis(float pos, float flow, float size, float rsize)
{
    const float X = __fadd_rn(pos, flow);
    // g = 2 (X/size - 0.5) and (g + 1) / 2: doubling and halving are exact and commute with rounding, so
    // RN(RN(2 s) + 1) / 2 == RN(s + 0.5) bit for bit (s = RN(X/size - 0.5); no overflow or underflow here)
    const float s = __fsub_rn(div_by_size(X, size, rsize), 0.5f);
    return __fmul_rn(__fadd_rn(s, 0.5f), size - 1.f);
}

struct WarpGeom {
    int W, H;
    float fW, fH, rW, rH;  // rW = RN(1 / W), rH = RN(1 / H) (host, IEEE division)
};

static inline WarpGeom warp_geom(int H, int W)
{
    WarpGeom g;
    g.W = W; g.H = H;
    g.fW = (float)W; g.fH = (float)H;
    g.rW = 1.0f / (float)W; g.rH = 1.0f / (float)H;
    return g;
}

__device__ __forceinline__ WarpCoord warp_coord(int x, int y, float u, float v, const WarpGeom &g)
{
    WarpCoord c;
    c.ix = warp_axis((float)x, u, g.fW, g.rW);
    c.iy = warp_axis((float)y, v, g.fH, g.rH);
    c.x0 = __float2int_rd(c.ix);
    c.y0 = __float2int_rd(c.iy);
    return c;
}

// The four taps of one sample point: offset of the north-west tap, validity of each tap (zero padding) and the
// bilinear weights -- formed once per pixel and reused for every channel.
struct Taps {
    int o;
    bool v00, v01, v10, v11;
    float wnw, wne, wsw, wse;
};

__device__ __forceinline__ Taps make_taps(const WarpCoord &c, int H, int W)
{
    Taps t;
    const float x1 = (float)(c.x0 + 1), y1 = (float)(c.y0 + 1), x0 = (float)c.x0, y0 = (float)c.y0;
    t.wnw = (x1 - c.ix) * (y1 - c.iy);
    t.wne = (c.ix - x0) * (y1 - c.iy);
    t.wsw = (x1 - c.ix) * (c.iy - y0);
    t.wse = (c.ix - x0) * (c.iy - y0);
    const bool vx0 = (unsigned)c.x0 < (unsigned)W, vx1 = (unsigned)c.x0 + 1u < (unsigned)W;
    const bool vy0 = (unsigned)c.y0 < (unsigned)H, vy1 = (unsigned)c.y0 + 1u < (unsigned)H;
    t.v00 = vx0 && vy0; t.v01 = vx1 && vy0; t.v10 = vx0 && vy1; t.v11 = vx1 && vy1;
    // the offset is only dereferenced where a tap is valid; clamping keeps the product inside int range
    t.o = min(max(c.y0, -1), H) * W + min(max(c.x0, -1), W);
    return t;
}

__device__ __forceinline__ float sample(const float *__restrict__ plane, const Taps &t, int W)
{
    const float *p = plane + t.o;
    const float nw = t.v00 ? __ldg(p) : 0.f, ne = t.v01 ? __ldg(p + 1) : 0.f;
    const float sw = t.v10 ? __ldg(p + W) : 0.f, se = t.v11 ? __ldg(p + W + 1) : 0.f;
    return nw * t.wnw + ne * t.wne + sw * t.wsw + se * t.wse;
}

// C == 0: runtime channel count
#define TAI_CH_LOOP(CT, C) _Pragma("unroll") for (int ch = 0; ch < ((CT) ? (CT) : (C)); ++ch)

}  // namespace tai
