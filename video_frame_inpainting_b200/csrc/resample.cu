// HBM-bound resampling kernels of the TAI / MC-Net decoders for sm_100a:
//
//   * bilinear x2 upsample with the torch-0.3.1 mapping (today's align_corners=True), forward and its
//     adjoint -- nn.Upsample(scale_factor=2, mode='bilinear') in the kernel-net decoder and heads
//     (tai.py:283,337,343) and in the Super-SloMo decoders (slomo.py:113-149).  The library kernel behind
//     nn.Upsample writes one element per thread and was 5.5 % of the KTH training step
//     (profiles/r01_step_kernels.csv); here a thread writes four consecutive outputs with one 128-bit store;
//   * zero-insertion unpooling fused with the residual add of DecCnn (mcnet.py:234-236, 240-256: two cats,
//     a clone().zero_(), two permutes and an add in the reference), forward and adjoint.
//
// All four are pure streaming kernels: the forward kernels are bound by the write of the 2H x 2W result.
#include "common.cuh"

namespace tai {

static inline unsigned resample_grid(long work_items, int block)
{
    // grid-stride over 16 CTAs per SM (measured: an uncapped one-item-per-thread grid is 7-15 % slower here)
    long g = (work_items + block - 1) / block;
    const long cap = (long)sm_count() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

// Source coordinate of destination index d, exactly as the reference's library computed it for the
// align-corners mapping: src = d * (in - 1) / (out - 1) in FP32, i0 = (int)src, lambda = src - i0,
// i1 = i0 + (i0 < in - 1).
struct Tap2 {
    int i0, i1;
    float w0, w1;
};
__device__ __forceinline__ Tap2 up_tap(int d, float ratio, int in)
{
    Tap2 t;
    const float src = ratio * (float)d;
    t.i0 = (int)src;
    t.i1 = t.i0 + (t.i0 < in - 1 ? 1 : 0);
    t.w1 = src - (float)t.i0;
    t.w0 = 1.f - t.w1;
    return t;
}

// ---- forward ---------------------------------------------------------------------------------------
// out[n, oy, ox] = w0y * (w0x * I[y0,x0] + w1x * I[y0,x1]) + w1y * (w0x * I[y1,x0] + w1x * I[y1,x1])
//
// Separable, through shared memory: a CTA owns a UF_TH x UF_TW output tile; it stages the (at most)
// UF_TH/2+2 input rows x UF_TW/2+2 input columns the tile touches, interpolates them horizontally into a
// row buffer (every input row is interpolated once, not once per output row that uses it) and then
// vertically into the result.  Each input element is loaded from global memory once per tile and each
// output costs ~2 shared-memory reads; the first version (four global loads per output element) was bound
// by LSU issue at 42 % of the copy bandwidth.
constexpr int UF_TH = 32, UF_TW = 128, UF_NT = 256;
constexpr int UF_IR = UF_TH / 2 + 3, UF_IC = UF_TW / 2 + 3;

template <bool VEC4>
__global__ void __launch_bounds__(UF_NT)
upsample2x_fwd_kernel(const float *__restrict__ in, float *__restrict__ out, int H, int W, float rh, float rw, int tiles_y,
                      int tiles_x)
{
    __shared__ float s_in[UF_IR][UF_IC + 1];
    __shared__ __align__(16) float s_row[UF_IR][UF_TW];
    const int Ho = 2 * H, Wo = 2 * W;
    int t = blockIdx.x;
    const int tx = t % tiles_x;
    t /= tiles_x;
    const int ty = t % tiles_y;
    const long n = t / tiles_y;
    const int oy0 = ty * UF_TH, ox0 = tx * UF_TW;
    const int oy1 = min(oy0 + UF_TH, Ho) - 1, ox1 = min(ox0 + UF_TW, Wo) - 1;   // last output row / column of the tile
    const int r_lo = up_tap(oy0, rh, H).i0, r_hi = up_tap(oy1, rh, H).i1;
    const int c_lo = up_tap(ox0, rw, W).i0, c_hi = up_tap(ox1, rw, W).i1;
    const int nr = r_hi - r_lo + 1, nc = c_hi - c_lo + 1;                       // <= UF_IR, UF_IC
    {   // stage the input tile: 64 threads per row (two passes cover the <= 67 columns), 4 rows at a time
        const float *src = in + (n * H + r_lo) * W + c_lo;
        const int c = threadIdx.x & 63;
        constexpr int NIT = (UF_IR + UF_NT / 64 - 1) / (UF_NT / 64);
        float a[NIT], b[NIT];
#pragma unroll
        for (int k = 0; k < NIT; ++k) {  // every load is in flight before the first store
            const int r = (threadIdx.x >> 6) + k * (UF_NT / 64);
            a[k] = (r < nr && c < nc) ? __ldg(src + (long)r * W + c) : 0.f;
            b[k] = (r < nr && c + 64 < nc) ? __ldg(src + (long)r * W + c + 64) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < NIT; ++k) {
            const int r = (threadIdx.x >> 6) + k * (UF_NT / 64);
            if (r < nr) {
                s_in[r][c] = a[k];
                if (c + 64 < UF_IC + 1) s_in[r][c + 64] = b[k];
            }
        }
    }
    __syncthreads();
    {   // horizontal pass: thread = one output column (its taps are formed once), loops over the input rows
        const int oxl = threadIdx.x % UF_TW;
        const Tap2 tp = up_tap(min(ox0 + oxl, ox1), rw, W);
        const int a0 = tp.i0 - c_lo, a1 = tp.i1 - c_lo;
#pragma unroll 3
        for (int r = threadIdx.x / UF_TW; r < nr; r += UF_NT / UF_TW)
            s_row[r][oxl] = tp.w0 * s_in[r][a0] + tp.w1 * s_in[r][a1];
    }
    __syncthreads();
    if (VEC4) {  // vertical pass, four consecutive output columns per thread: one tap evaluation and one 128-bit store per 4
        const int q = threadIdx.x & 31;
        if (ox0 + 4 * q <= ox1) {
            float *obase = out + n * Ho * Wo + ox0 + 4 * q;
#pragma unroll 2
            for (int oy = oy0 + (threadIdx.x >> 5); oy <= oy1; oy += UF_NT / 32) {
                const Tap2 tp = up_tap(oy, rh, H);
                const float4 a = *reinterpret_cast<const float4 *>(&s_row[tp.i0 - r_lo][4 * q]);
                const float4 b = *reinterpret_cast<const float4 *>(&s_row[tp.i1 - r_lo][4 * q]);
                float4 o;
                o.x = tp.w0 * a.x + tp.w1 * b.x;
                o.y = tp.w0 * a.y + tp.w1 * b.y;
                o.z = tp.w0 * a.z + tp.w1 * b.z;
                o.w = tp.w0 * a.w + tp.w1 * b.w;
                *reinterpret_cast<float4 *>(obase + (long)oy * Wo) = o;
            }
        }
    } else {     // any width / alignment: consecutive threads write consecutive output columns
        const int oxl = threadIdx.x % UF_TW;
        if (ox0 + oxl <= ox1) {
            float *obase = out + n * Ho * Wo + ox0 + oxl;
            for (int oy = oy0 + threadIdx.x / UF_TW; oy <= oy1; oy += UF_NT / UF_TW) {
                const Tap2 tp = up_tap(oy, rh, H);
                obase[(long)oy * Wo] = tp.w0 * s_row[tp.i0 - r_lo][oxl] + tp.w1 * s_row[tp.i1 - r_lo][oxl];
            }
        }
    }
}

// Register sliding-window forward (used below W = 64, see the launcher; needs an even W so that the 2W-wide result
// rows take 128-bit stores): a thread owns FOUR consecutive output columns of one plane and walks R output rows downwards.  The four
// columns touch at most the input columns cb .. cb+3; their horizontal interpolation is a 4 x 4 weight matrix
// (two non-zeros per row, formed once per thread), applied once per INPUT row (16 FMAs); an output row is one
// vertical blend of the two interpolated input rows in registers and one 128-bit store.  The next input row is
// loaded one step ahead.  ~6 instructions per output element: the shared-memory tile version (stage, horizontal
// pass, vertical pass, two barriers) spent ~30 and was ISSUE-bound (ncu: issue slots 85 % busy, DRAM at 35 %).
constexpr int UR_NT = 256;

__global__ void __launch_bounds__(UR_NT)
upsample2x_fwd_reg_kernel(const float *__restrict__ in, float *__restrict__ out, long planes, int H, int W, float rh, float rw,
                          int R)
{
    const int Ho = 2 * H, Wo = 2 * W, wq = Wo / 4;
    const int nrb = (Ho + R - 1) / R;
    const int oy_begin = (int)(blockIdx.x % nrb) * R;       // row block fastest: CTAs sharing input rows run together
    const long item = (long)(blockIdx.x / nrb) * UR_NT + threadIdx.x;
    if (item >= planes * wq) return;
    const long n = item / wq;
    const int ox0 = 4 * (int)(item - n * wq);
    const int cb = min(up_tap(ox0, rw, W).i0, max(W - 4, 0));
    float M[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const Tap2 t = up_tap(ox0 + k, rw, W);
#pragma unroll
        for (int j = 0; j < 4; ++j) M[k][j] = (t.i0 == cb + j ? t.w0 : 0.f) + (t.i1 == cb + j ? t.w1 : 0.f);
    }
    int col[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) col[j] = min(cb + j, W - 1);
    const float *p = in + n * H * W;
    float *o = out + n * Ho * Wo + ox0;
    auto load_row = [&](int r, float (&v)[4]) {
        const float *row = p + min(r, H - 1) * W;
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __ldg(row + col[j]);
    };
    auto interp = [&](const float (&v)[4], float (&h)[4]) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float a = M[k][0] * v[0];
            a = fmaf(M[k][1], v[1], a);
            a = fmaf(M[k][2], v[2], a);
            h[k] = fmaf(M[k][3], v[3], a);
        }
    };
    int lo = up_tap(oy_begin, rh, H).i0;   // input row held in h0; h1 holds row min(lo + 1, H - 1)
    float h0[4], h1[4], v[4], vn[4];
    load_row(lo, v);
    interp(v, h0);
    load_row(lo + 1, v);
    interp(v, h1);
    load_row(lo + 2, vn);                  // one input row ahead
    const int oy_end = min(Ho, oy_begin + R);
    for (int oy = oy_begin; oy < oy_end; ++oy) {
        const Tap2 t = up_tap(oy, rh, H);
        if (t.i0 != lo) {                  // uniform over the CTA; the ratio is < 1/2, so i0 advances by at most one
            lo = t.i0;
#pragma unroll
            for (int k = 0; k < 4; ++k) h0[k] = h1[k];
            interp(vn, h1);
            load_row(lo + 2, vn);
        }
        float4 r;
        r.x = t.w0 * h0[0] + t.w1 * h1[0];
        r.y = t.w0 * h0[1] + t.w1 * h1[1];
        r.z = t.w0 * h0[2] + t.w1 * h1[2];
        r.w = t.w0 * h0[3] + t.w1 * h1[3];
        *reinterpret_cast<float4 *>(o + (long)oy * Wo) = r;
    }
}

// ---- adjoint ----------------------------------------------------------------------------------------
// Evaluated as a gather (deterministic, no atomics): input pixel (y, x) collects every output pixel whose
// two taps per axis include it.  With ratio = (in-1)/(2in-1) < 1/2 those are among 2y-2 .. 2y+3.
// weight of output index d on input index i along one axis
__device__ __forceinline__ float up_adjoint_weight(int d, int i, float ratio, int in, int out)
{
    float acc = 0.f;
    if (d >= 0 && d < out) {
        const Tap2 t = up_tap(d, ratio, in);
        if (t.i0 == i) acc += t.w0;
        if (t.i1 == i) acc += t.w1;
    }
    return acc;
}

// Register sliding window, no shared-memory staging: a thread owns the input-gradient columns (x, x+1), x even,
// of one plane and walks R input rows downwards.  For every output-gradient row d it loads eight consecutive
// columns starting at cb = clamp(2x-2, 0, 2W-8) (four 64-bit loads off ONE row pointer with immediate offsets;
// the neighbouring threads' overlap is served by L1), reduces them along x with two sets of eight column
// weights (registers, formed once from the columns' real indices, so the shifted window of the two border
// threads needs no special case and no load is ever out of bounds) and keeps the values of the six rows
// 2y-2 .. 2y+3 in a register window that slides by two rows per input row; the six row weights of an input row
// are the same for the whole CTA and come from a small shared-memory table (broadcast reads).
// Per input element: 4 LDG.64, 22 FMAs, half a 64-bit store -- the separable shared-memory version (stage,
// reduce along x, reduce along y: ~20 shared/global instructions per input element, two barriers per tile) ran
// at 23-45 % of the copy bandwidth, bound by LSU issue.  Deterministic (a gather, no atomics).
constexpr int UA_NT = 256, UA_RMAX = 32;

// rows outside the image are read from a clamped address and replaced by zeros afterwards (a select, so that
// an Inf in the border row cannot turn into 0 * Inf); no branches: every load of a step is in flight before
// the first use
__device__ __forceinline__ void up_adj_load_row(const float *__restrict__ pc, int d, int Ho, int Wo, float2 (&v)[4])
{
    const float2 *rp = reinterpret_cast<const float2 *>(pc + min(max(d, 0), Ho - 1) * Wo);
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __ldg(rp + k);
}

__device__ __forceinline__ void up_adj_reduce_x(const float2 (&v)[4], const float (&wx0)[8], const float (&wx1)[8], bool rok,
                                                float &t0, float &t1)
{
    float a = wx0[0] * v[0].x, b = wx1[0] * v[0].x;
    a = fmaf(wx0[1], v[0].y, a); b = fmaf(wx1[1], v[0].y, b);
#pragma unroll
    for (int k = 1; k < 4; ++k) {
        a = fmaf(wx0[2 * k], v[k].x, a); b = fmaf(wx1[2 * k], v[k].x, b);
        a = fmaf(wx0[2 * k + 1], v[k].y, a); b = fmaf(wx1[2 * k + 1], v[k].y, b);
    }
    t0 = rok ? a : 0.f;
    t1 = rok ? b : 0.f;
}

// needs W even, W >= 4 and 8 B-aligned tensors
__global__ void __launch_bounds__(UA_NT)
upsample2x_bwd_kernel(const float *__restrict__ gout, float *__restrict__ gin, long planes, int H, int W, float rh, float rw,
                      int R)
{
    __shared__ __align__(16) float s_wy[UA_RMAX][8];
    const int Ho = 2 * H, Wo = 2 * W, wp = W / 2;
    // the row block is the fastest-varying part of the CTA index: the CTAs that share halo rows run together
    const int nrb = (H + R - 1) / R;
    const int y_begin = (int)(blockIdx.x % nrb) * R;
    const long bx = blockIdx.x / nrb;
    for (int i = threadIdx.x; i < R * 6; i += UA_NT) {  // row weights of this CTA's R input rows
        const int e = i / 6, k = i - e * 6;
        const int y = y_begin + e;
        s_wy[e][k] = (y < H) ? up_adjoint_weight(2 * y - 2 + k, y, rh, H, Ho) : 0.f;
    }
    __syncthreads();
    const long item = bx * UA_NT + threadIdx.x;
    if (item >= planes * wp) return;
    const long n = item / wp;
    const int x = 2 * (int)(item - n * wp);
    const int cb = min(max(2 * x - 2, 0), Wo - 8);
    float wx0[8], wx1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        wx0[k] = up_adjoint_weight(cb + k, x, rw, W, Wo);
        wx1[k] = up_adjoint_weight(cb + k, x + 1, rw, W, Wo);
    }
    const float *pc = gout + n * Ho * Wo + cb;
    float *o = gin + n * H * W + x;

    float a0[4], a1[4];  // x-reduced rows 2y-2 .. 2y+1 of the current input row y
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int d = 2 * y_begin - 2 + r;
        float2 v[4];
        up_adj_load_row(pc, d, Ho, Wo, v);
        up_adj_reduce_x(v, wx0, wx1, d >= 0 && d < Ho, a0[r], a1[r]);
    }
    const int y_end = min(H, y_begin + R);
    // two input rows per step = four new gradient rows, all sixteen loads first.  (An explicitly software-pipelined
    // form -- the loads of step s+1 issued before the arithmetic of step s, 80 registers, 3 CTAs/SM -- measured
    // 49 / 33 / 53 us against 46 / 29 / 55 us for this one (48 registers, 5 CTAs/SM) at [32,64,64,64] /
    // [32,128,32,32] / [8,64,120,160]; rows per thread 8 / 16 / 32 are within 10 % of each other.)
#pragma unroll 1
    for (int y = y_begin; y < y_end; y += 2) {
        float2 v[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) up_adj_load_row(pc, 2 * y + 2 + r, Ho, Wo, v[r]);
        float t0[4], t1[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) up_adj_reduce_x(v[r], wx0, wx1, 2 * y + 2 + r < Ho, t0[r], t1[r]);
        const float4 wa = *reinterpret_cast<const float4 *>(&s_wy[y - y_begin][0]);
        const float2 wb = *reinterpret_cast<const float2 *>(&s_wy[y - y_begin][4]);
        const float4 wc = *reinterpret_cast<const float4 *>(&s_wy[y - y_begin + 1][0]);
        const float2 wd = *reinterpret_cast<const float2 *>(&s_wy[y - y_begin + 1][4]);
        float p0 = wa.x * a0[0], p1 = wa.x * a1[0];
        p0 = fmaf(wa.y, a0[1], p0); p1 = fmaf(wa.y, a1[1], p1);
        p0 = fmaf(wa.z, a0[2], p0); p1 = fmaf(wa.z, a1[2], p1);
        p0 = fmaf(wa.w, a0[3], p0); p1 = fmaf(wa.w, a1[3], p1);
        p0 = fmaf(wb.x, t0[0], p0); p1 = fmaf(wb.x, t1[0], p1);
        p0 = fmaf(wb.y, t0[1], p0); p1 = fmaf(wb.y, t1[1], p1);
        float q0 = wc.x * a0[2], q1 = wc.x * a1[2];
        q0 = fmaf(wc.y, a0[3], q0); q1 = fmaf(wc.y, a1[3], q1);
        q0 = fmaf(wc.z, t0[0], q0); q1 = fmaf(wc.z, t1[0], q1);
        q0 = fmaf(wc.w, t0[1], q0); q1 = fmaf(wc.w, t1[1], q1);
        q0 = fmaf(wd.x, t0[2], q0); q1 = fmaf(wd.x, t1[2], q1);
        q0 = fmaf(wd.y, t0[3], q0); q1 = fmaf(wd.y, t1[3], q1);
        float *oy = o + y * W;
        *reinterpret_cast<float2 *>(oy) = make_float2(p0, p1);
        if (y + 1 < y_end) *reinterpret_cast<float2 *>(oy + W) = make_float2(q0, q1);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            a0[r] = t0[r];
            a1[r] = t1[r];
        }
    }
}

// Any shape / alignment (odd W, W < 4, unaligned views): one thread per input element gathers its 6 x 6
// neighbourhood directly.  Not a performance path.
__global__ void __launch_bounds__(256)
upsample2x_bwd_generic_kernel(const float *__restrict__ gout, float *__restrict__ gin, long total, int H, int W, float rh,
                              float rw)
{
    const int Ho = 2 * H, Wo = 2 * W;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % W);
        const long row = idx / W;
        const int y = (int)(row % H);
        const float *g = gout + (row / H) * Ho * Wo;
        float acc = 0.f;
        for (int i = 0; i < 6; ++i) {
            const int d = 2 * y - 2 + i;
            const float wy = up_adjoint_weight(d, y, rh, H, Ho);
            if (wy == 0.f) continue;
            float t = 0.f;
            for (int k = 0; k < 6; ++k) {
                const int e = 2 * x - 2 + k;
                const float wx = up_adjoint_weight(e, x, rw, W, Wo);
                if (wx != 0.f) t = fmaf(wx, g[(long)d * Wo + e], t);
            }
            acc = fmaf(wy, t, acc);
        }
        gin[idx] = acc;
    }
}

// out[n, 2y+dy, 2x+dx] = res[n, 2y+dy, 2x+dx] + (dy == 0 && dx == 0 ? x[n, y, x] : 0)
template <int VEC>
__global__ void __launch_bounds__(256)
unpool_add_fwd_kernel(const float *__restrict__ x, const float *__restrict__ res, float *__restrict__ out, long N, int H,
                      int W)
{
    const int Ho = 2 * H, Wo = 2 * W;
    const int wq = Wo / VEC;
    const long total = N * Ho * wq;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int q = (int)(idx % wq);
        const long row = idx / wq;
        const int oy = (int)(row % Ho);
        const long n = row / Ho;
        const long o = row * Wo + (long)q * VEC;
        if (VEC == 4) {
            float4 r = ld_stream4(reinterpret_cast<const float4 *>(res + o));
            if ((oy & 1) == 0) {
                const float2 v = *reinterpret_cast<const float2 *>(x + (n * H + (oy >> 1)) * W + 2 * q);
                r.x += v.x;
                r.z += v.y;
            }
            *reinterpret_cast<float4 *>(out + o) = r;
        } else {
            float r = res[o];
            if (((oy | q) & 1) == 0) r += x[(n * H + (oy >> 1)) * W + (q >> 1)];
            out[o] = r;
        }
    }
}

// g_x[n, y, x] = g_out[n, 2y, 2x]   (the residual's gradient is g_out itself)
template <int VEC>
__global__ void __launch_bounds__(256)
unpool_bwd_kernel(const float *__restrict__ gout, float *__restrict__ gx, long N, int H, int W)
{
    const int Wo = 2 * W;
    const int wq = W / VEC;
    const long total = N * H * wq;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int q = (int)(idx % wq);
        const long row = idx / wq;  // n * H + y
        const float *g = gout + (2 * row) * Wo + (long)q * 2 * VEC;
        if (VEC == 2) {
            const float4 v = ld_stream4(reinterpret_cast<const float4 *>(g));
            *reinterpret_cast<float2 *>(gx + row * W + 2 * q) = make_float2(v.x, v.z);
        } else {
            gx[row * W + q] = g[0];
        }
    }
}

// ---- 2 x 2 max pooling (nn.MaxPool2d(2): mcnet.py:28-45, slomo.py:47-85) ------------------------------
// out[n,y,x] = max of in[n, 2y..2y+1, 2x..2x+1]; the position of the maximum is kept as a 2-bit code (one byte
// per output element) instead of the library's int64 flat index: the backward kernel reads 1 B instead of 8 B
// per output element and writes every input-gradient element exactly once (no zero-fill + scatter).
// Selection rule of the library kernel (max_pool_forward_nchw): scan (0,0), (0,1), (1,0), (1,1), take a later
// element only if it is GREATER or NaN -- so ties (the zeros behind a ReLU) go to the first position.  The
// code is the integer part of the op and is compared bit-exactly in the tests.  Odd H / W: floor mode, the
// last row / column is not pooled and receives a zero gradient.
__device__ __forceinline__ void mp_pick(float v, int k, float &m, int &code)
{
    if (v > m || v != v) {
        m = v;
        code = k;
    }
}

template <bool VEC>
__global__ void __launch_bounds__(256)
maxpool2x2_fwd_kernel(const float *__restrict__ in, float *__restrict__ out, unsigned char *__restrict__ code, long N, int H,
                      int W)
{
    const int Ho = H / 2, Wo = W / 2;
    const int wq = VEC ? Wo / 2 : Wo;  // VEC: two outputs (four input columns) per thread
    const long total = N * Ho * wq;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int q = (int)(idx % wq);
        const long row = idx / wq;  // n * Ho + y
        const int y = (int)(row % Ho);
        const long n = row / Ho;
        const float *p = in + (n * H + 2 * y) * W;
        if (VEC) {
            const float4 a = ld_stream4(reinterpret_cast<const float4 *>(p + 4 * q));
            const float4 b = ld_stream4(reinterpret_cast<const float4 *>(p + W + 4 * q));
            float m0 = a.x, m1 = a.z;
            int c0 = 0, c1 = 0;
            mp_pick(a.y, 1, m0, c0); mp_pick(b.x, 2, m0, c0); mp_pick(b.y, 3, m0, c0);
            mp_pick(a.w, 1, m1, c1); mp_pick(b.z, 2, m1, c1); mp_pick(b.w, 3, m1, c1);
            *reinterpret_cast<float2 *>(out + row * Wo + 2 * q) = make_float2(m0, m1);
            *reinterpret_cast<uchar2 *>(code + row * Wo + 2 * q) = make_uchar2((unsigned char)c0, (unsigned char)c1);
        } else {
            float m = p[2 * q];
            int c = 0;
            mp_pick(p[2 * q + 1], 1, m, c); mp_pick(p[W + 2 * q], 2, m, c); mp_pick(p[W + 2 * q + 1], 3, m, c);
            out[row * Wo + q] = m;
            code[row * Wo + q] = (unsigned char)c;
        }
    }
}

// gin[n, 2y+dy, 2x+dx] = (code[n,y,x] == 2*dy+dx) ? gout[n,y,x] : 0; unpooled last row / column: 0
template <bool VEC>
__global__ void __launch_bounds__(256)
maxpool2x2_bwd_kernel(const float *__restrict__ gout, const unsigned char *__restrict__ code, float *__restrict__ gin, long N,
                      int H, int W)
{
    const int Ho = H / 2, Wo = W / 2;
    if (VEC) {  // H, W even here: every input element belongs to a window
        const int wq = Wo / 2;
        const long total = N * Ho * wq;
        for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
            const int q = (int)(idx % wq);
            const long row = idx / wq;
            const int y = (int)(row % Ho);
            const long n = row / Ho;
            const float2 g = *reinterpret_cast<const float2 *>(gout + row * Wo + 2 * q);
            const uchar2 c = *reinterpret_cast<const uchar2 *>(code + row * Wo + 2 * q);
            float *p = gin + (n * H + 2 * y) * W + 4 * q;
            *reinterpret_cast<float4 *>(p) = make_float4(c.x == 0 ? g.x : 0.f, c.x == 1 ? g.x : 0.f, c.y == 0 ? g.y : 0.f,
                                                         c.y == 1 ? g.y : 0.f);
            *reinterpret_cast<float4 *>(p + W) = make_float4(c.x == 2 ? g.x : 0.f, c.x == 3 ? g.x : 0.f, c.y == 2 ? g.y : 0.f,
                                                             c.y == 3 ? g.y : 0.f);
        }
    } else {    // one thread per INPUT element
        const long total = N * H * W;
        for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
            const int x = (int)(idx % W);
            const long row = idx / W;
            const int y = (int)(row % H);
            const long n = row / H;
            float v = 0.f;
            if ((y >> 1) < Ho && (x >> 1) < Wo) {
                const long o = (n * Ho + (y >> 1)) * Wo + (x >> 1);
                if (code[o] == (unsigned char)(2 * (y & 1) + (x & 1))) v = gout[o];
            }
            gin[idx] = v;
        }
    }
}

// `scale`: elements of the LARGEST tensor of the call per H x W plane element (4 where H x W is the low-resolution
// side of an up / down-sampling pair, 1 where it is the high-resolution side)
static int resample_args_ok(const char *who, const void *a, const void *b, long long N, int H, int W, int scale = 4)
{
    TAI_REQUIRE(a && b && N > 0 && H > 0 && W > 0, TAI_ERR_INVALID_ARGUMENT, "%s: bad arguments N=%lld H=%d W=%d", who, N, H, W);
    TAI_REQUIRE(fits_int31(N * (long long)scale * H * W), TAI_ERR_TOO_LARGE, "%s: tensor has >= 2^31 elements", who);
    return TAI_OK;
}

static inline bool aligned16(const void *a, const void *b, const void *c = nullptr)
{
    return ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0);
}

}  // namespace tai

using namespace tai;

extern "C" int upsample_bilinear2x_forward_b200(const float *in, float *out, long long N, int H, int W, void *stream)
{
    int rc = resample_args_ok("upsample_bilinear2x_forward_b200", in, out, N, H, W);
    if (rc != TAI_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // the library computed the ratio in FP32: (in - 1) / (out - 1)
    const float rh = (float)(H - 1) / (float)(2 * H - 1), rw = (float)(W - 1) / (float)(2 * W - 1);
    const long out_el = (long)N * 4 * H * W;
    const int tiles_y = ceil_div(2 * H, UF_TH), tiles_x = ceil_div(2 * W, UF_TW);
    const long long blocks = N * tiles_y * tiles_x;
    TAI_REQUIRE(blocks < (1LL << 31), TAI_ERR_TOO_LARGE, "upsample_bilinear2x_forward_b200: too many tiles");
    TimingScope ts("upsample2x_fwd", st, 0.0, 4.0 * (out_el + out_el / 4));  // read the input once, write the result
    // measured (B200, us): [32,64,64,64] tile 38.9 / register 49.2; [32,32,64,64] 23.3 / 28.8; [8,64,120,160] 55.1 / 53.3;
    // [32,128,32,32] 38.9 / 29.7 -- the 128-column tile wastes half its threads below W = 64, the register walk is
    // latency-bound on large planes: pick by width
    if ((W % 2) == 0 && W < 64 && aligned16(out, out)) {
        const long planes = (long)N;
        const int wq = W / 2;                  // threads per output row
        int R = 32;                            // output rows per thread; shorter walks when the tensor would not fill the chip
        while (R > 8 && planes * wq * ceil_div(2 * H, R) < (long)sm_count() * 1024) R /= 2;
        const long long bx = ((planes * wq + UR_NT - 1) / UR_NT) * ceil_div(2 * H, R);
        TAI_REQUIRE(bx < (1LL << 31), TAI_ERR_TOO_LARGE, "upsample_bilinear2x_forward_b200: grid too large");
        upsample2x_fwd_reg_kernel<<<(unsigned)bx, UR_NT, 0, st>>>(in, out, planes, H, W, rh, rw, R);
    } else if ((W % 2) == 0 && aligned16(out, out)) {
        upsample2x_fwd_kernel<true><<<(unsigned)blocks, UF_NT, 0, st>>>(in, out, H, W, rh, rw, tiles_y, tiles_x);
    } else {
        upsample2x_fwd_kernel<false><<<(unsigned)blocks, UF_NT, 0, st>>>(in, out, H, W, rh, rw, tiles_y, tiles_x);
    }
    return check_launch("upsample2x_fwd_kernel");
}

extern "C" int upsample_bilinear2x_backward_b200(const float *grad_out, float *grad_in, long long N, int H, int W,
                                                 void *stream)
{
    int rc = resample_args_ok("upsample_bilinear2x_backward_b200", grad_out, grad_in, N, H, W);
    if (rc != TAI_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const float rh = (float)(H - 1) / (float)(2 * H - 1), rw = (float)(W - 1) / (float)(2 * W - 1);
    const long in_el = (long)N * H * W;
    TimingScope ts("upsample2x_bwd", st, 0.0, 4.0 * (5 * in_el));  // read the 2H x 2W gradient once, write H x W
    if ((W % 2) == 0 && W >= 4 && (((uintptr_t)grad_out | (uintptr_t)grad_in) & 7) == 0) {
        const long planes = (long)N;
        const int wp = W / 2;
        int R = 16;  // input rows per thread (even); shorter walks when the tensor would not fill the chip
        while (R > 4 && planes * wp * ceil_div(H, R) < (long)sm_count() * 1024) R /= 2;
        const long long bx = ((planes * wp + UA_NT - 1) / UA_NT) * ceil_div(H, R);
        TAI_REQUIRE(bx < (1LL << 31), TAI_ERR_TOO_LARGE, "upsample_bilinear2x_backward_b200: grid too large");
        const unsigned grid = (unsigned)bx;
        upsample2x_bwd_kernel<<<grid, UA_NT, 0, st>>>(grad_out, grad_in, planes, H, W, rh, rw, R);
    } else {
        upsample2x_bwd_generic_kernel<<<resample_grid(in_el, 256), 256, 0, st>>>(grad_out, grad_in, in_el, H, W, rh, rw);
    }
    return check_launch("upsample2x_bwd_kernel");
}

extern "C" int unpool_add_forward_b200(const float *x, const float *res, float *out, long long N, int H, int W,
                                       void *stream)
{
    int rc = resample_args_ok("unpool_add_forward_b200", x, out, N, H, W);
    if (rc != TAI_OK) return rc;
    TAI_REQUIRE(res != nullptr, TAI_ERR_INVALID_ARGUMENT, "unpool_add_forward_b200: null residual");
    cudaStream_t st = (cudaStream_t)stream;
    const long out_el = (long)N * 4 * H * W;
    TimingScope ts("unpool_add_fwd", st, 0.0, 4.0 * (2 * out_el + out_el / 4));
    if ((W % 2) == 0 && aligned16(res, out) && ((uintptr_t)x & 7) == 0)
        unpool_add_fwd_kernel<4><<<resample_grid(out_el / 4, 256), 256, 0, st>>>(x, res, out, (long)N, H, W);
    else
        unpool_add_fwd_kernel<1><<<resample_grid(out_el, 256), 256, 0, st>>>(x, res, out, (long)N, H, W);
    return check_launch("unpool_add_fwd_kernel");
}

extern "C" int unpool_backward_b200(const float *grad_out, float *grad_x, long long N, int H, int W, void *stream)
{
    int rc = resample_args_ok("unpool_backward_b200", grad_out, grad_x, N, H, W);
    if (rc != TAI_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const long in_el = (long)N * H * W;
    TimingScope ts("unpool_bwd", st, 0.0, 4.0 * (2 * in_el));
    if ((W % 2) == 0 && aligned16(grad_out, grad_out) && ((uintptr_t)grad_x & 7) == 0)
        unpool_bwd_kernel<2><<<resample_grid(in_el / 2, 256), 256, 0, st>>>(grad_out, grad_x, (long)N, H, W);
    else
        unpool_bwd_kernel<1><<<resample_grid(in_el, 256), 256, 0, st>>>(grad_out, grad_x, (long)N, H, W);
    return check_launch("unpool_bwd_kernel");
}

extern "C" int maxpool2x2_forward_b200(const float *in, float *out, unsigned char *code, long long N, int H, int W, void *stream)
{
    int rc = resample_args_ok("maxpool2x2_forward_b200", in, out, N, H, W, 1);
    if (rc != TAI_OK) return rc;
    TAI_REQUIRE(code != nullptr && H >= 2 && W >= 2, TAI_ERR_INVALID_ARGUMENT, "maxpool2x2_forward_b200: needs H, W >= 2 and a code buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const long out_el = (long)N * (H / 2) * (W / 2);
    TimingScope ts("maxpool2x2_fwd", st, 0.0, 4.0 * N * H * W + 5.0 * out_el);  // read the input, write values + codes
    if ((W % 4) == 0 && (H % 2) == 0 && aligned16(in, in) && (((uintptr_t)out & 7) == 0) && (((uintptr_t)code & 1) == 0))
        maxpool2x2_fwd_kernel<true><<<resample_grid(out_el / 2, 256), 256, 0, st>>>(in, out, code, (long)N, H, W);
    else
        maxpool2x2_fwd_kernel<false><<<resample_grid(out_el, 256), 256, 0, st>>>(in, out, code, (long)N, H, W);
    return check_launch("maxpool2x2_fwd_kernel");
}

extern "C" int maxpool2x2_backward_b200(const float *grad_out, const unsigned char *code, float *grad_in, long long N, int H, int W,
                                        void *stream)
{
    int rc = resample_args_ok("maxpool2x2_backward_b200", grad_out, grad_in, N, H, W, 1);
    if (rc != TAI_OK) return rc;
    TAI_REQUIRE(code != nullptr && H >= 2 && W >= 2, TAI_ERR_INVALID_ARGUMENT, "maxpool2x2_backward_b200: needs H, W >= 2 and a code buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const long out_el = (long)N * (H / 2) * (W / 2);
    TimingScope ts("maxpool2x2_bwd", st, 0.0, 4.0 * N * H * W + 5.0 * out_el);  // read gradients + codes, write the input gradient
    if ((W % 4) == 0 && (H % 2) == 0 && aligned16(grad_in, grad_in) && (((uintptr_t)grad_out & 7) == 0) && (((uintptr_t)code & 1) == 0))
        maxpool2x2_bwd_kernel<true><<<resample_grid(out_el / 2, 256), 256, 0, st>>>(grad_out, code, grad_in, (long)N, H, W);
    else
        maxpool2x2_bwd_kernel<false><<<resample_grid((long)N * H * W, 256), 256, 0, st>>>(grad_out, code, grad_in, (long)N, H, W);
    return check_launch("maxpool2x2_bwd_kernel");
}
