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

// Separable as well: a CTA owns a UB_TH x UB_TW tile of the INPUT gradient; it stages the output-gradient
// rows 2*y0-2 .. 2*(y0+UB_TH-1)+3 and columns 2*x0-2 .. 2*(x0+UB_TW-1)+3 once (zero outside the image),
// reduces them along x with the six column weights of each input column, then along y with the six row
// weights of each input row.  (The direct form -- every thread loading its 6 x 6 neighbourhood -- issued 12
// global loads per input element and ran at 23 % of the copy bandwidth.)
constexpr int UB_TH = 16, UB_TW = 64, UB_NT = 256;
constexpr int UB_GR = 2 * UB_TH + 4, UB_GC = 2 * UB_TW + 4;

__global__ void __launch_bounds__(UB_NT)
upsample2x_bwd_kernel(const float *__restrict__ gout, float *__restrict__ gin, int H, int W, float rh, float rw,
                      int tiles_y, int tiles_x)
{
    __shared__ __align__(8) float s_g[UB_GR][UB_GC + 2];  // even pitch: 64-bit reads of column pairs
    __shared__ float s_t[UB_GR][UB_TW + 8];                // rows 2 apart land in different bank halves
    __shared__ __align__(16) float s_w[UB_TW + UB_TH][8];  // six adjoint weights per input column, then per input row
    const int Ho = 2 * H, Wo = 2 * W;
    int t = blockIdx.x;
    const int tx = t % tiles_x;
    t /= tiles_x;
    const int ty = t % tiles_y;
    const long n = t / tiles_y;
    const int y0 = ty * UB_TH, x0 = tx * UB_TW;
    const int oyb = 2 * y0 - 2, oxb = 2 * x0 - 2;
    const float *g = gout + n * Ho * Wo;
    for (int i = threadIdx.x; i < (UB_TW + UB_TH) * 6; i += UB_NT) {  // weights: formed once per CTA
        const int e = i / 6, k = i - e * 6;
        float w;
        if (e < UB_TW) {
            const int x = x0 + e;
            w = (x < W) ? up_adjoint_weight(2 * x - 2 + k, x, rw, W, Wo) : 0.f;
        } else {
            const int y = y0 + e - UB_TW;
            w = (y < H) ? up_adjoint_weight(2 * y - 2 + k, y, rh, H, Ho) : 0.f;
        }
        s_w[e][k] = w;
    }
    {   // stage: 66 column PAIRS per row (64-bit loads when the pair is inside the image), 3 rows at a time
        const int cp = threadIdx.x % 66, rl = threadIdx.x / 66;  // rl == 3 for the last 58 threads: idle
        const int ox = oxb + 2 * cp;
        const bool pair_ok = (Wo % 2 == 0) && ox >= 0 && ox + 1 < Wo && ((reinterpret_cast<uintptr_t>(g) & 7) == 0);
        if (rl < 3) {
            // all twelve loads of a thread are issued before the first store (the loop form waited for each
            // load in turn: the kernel was bound by global-load latency, not bandwidth)
            float2 v[UB_GR / 3];
#pragma unroll
            for (int k = 0; k < UB_GR / 3; ++k) {
                const int oy = oyb + rl + 3 * k;
                v[k] = make_float2(0.f, 0.f);
                if (oy >= 0 && oy < Ho) {
                    const float *gp = g + (long)oy * Wo + ox;
                    if (pair_ok) {
                        v[k] = __ldg(reinterpret_cast<const float2 *>(gp));
                    } else {
                        if (ox >= 0 && ox < Wo) v[k].x = __ldg(gp);
                        if (ox + 1 >= 0 && ox + 1 < Wo) v[k].y = __ldg(gp + 1);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < UB_GR / 3; ++k) *reinterpret_cast<float2 *>(&s_g[rl + 3 * k][2 * cp]) = v[k];
        }
    }
    __syncthreads();
    {   // along x: thread = one input column, six weights in registers, loops over the staged rows
        const int xl = threadIdx.x % UB_TW;
        const float4 wa = *reinterpret_cast<const float4 *>(&s_w[xl][0]);
        const float2 wb = *reinterpret_cast<const float2 *>(&s_w[xl][4]);
#pragma unroll 3
        for (int r = threadIdx.x / UB_TW; r < UB_GR; r += UB_NT / UB_TW) {
            const float2 *gp = reinterpret_cast<const float2 *>(&s_g[r][2 * xl]);
            const float2 g0 = gp[0], g1 = gp[1], g2 = gp[2];
            float acc = wa.x * g0.x;
            acc = fmaf(wa.y, g0.y, acc);
            acc = fmaf(wa.z, g1.x, acc);
            acc = fmaf(wa.w, g1.y, acc);
            acc = fmaf(wb.x, g2.x, acc);
            acc = fmaf(wb.y, g2.y, acc);
            s_t[r][xl] = acc;
        }
    }
    __syncthreads();
    {   // along y: thread = one input row (six weights in registers), loops over the columns
        const int yl = threadIdx.x / (UB_NT / UB_TH);            // 16 threads per input row
        const int y = y0 + yl;
        if (y < H) {
            const float4 wa = *reinterpret_cast<const float4 *>(&s_w[UB_TW + yl][0]);
            const float2 wb = *reinterpret_cast<const float2 *>(&s_w[UB_TW + yl][4]);
            float *orow = gin + (n * H + y) * W;
#pragma unroll
            for (int k = 0; k < UB_TW / (UB_NT / UB_TH); ++k) {
                const int xl = (threadIdx.x % (UB_NT / UB_TH)) + k * (UB_NT / UB_TH);
                float acc = wa.x * s_t[2 * yl][xl];
                acc = fmaf(wa.y, s_t[2 * yl + 1][xl], acc);
                acc = fmaf(wa.z, s_t[2 * yl + 2][xl], acc);
                acc = fmaf(wa.w, s_t[2 * yl + 3][xl], acc);
                acc = fmaf(wb.x, s_t[2 * yl + 4][xl], acc);
                acc = fmaf(wb.y, s_t[2 * yl + 5][xl], acc);
                if (x0 + xl < W) orow[x0 + xl] = acc;
            }
        }
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

static int resample_args_ok(const char *who, const void *a, const void *b, long long N, int H, int W)
{
    TAI_REQUIRE(a && b && N > 0 && H > 0 && W > 0, TAI_ERR_INVALID_ARGUMENT, "%s: bad arguments N=%lld H=%d W=%d", who, N, H, W);
    TAI_REQUIRE(fits_int31(N * 4LL * H * W), TAI_ERR_TOO_LARGE, "%s: tensor has >= 2^31 elements", who);
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
    if ((W % 2) == 0 && aligned16(out, out))
        upsample2x_fwd_kernel<true><<<(unsigned)blocks, UF_NT, 0, st>>>(in, out, H, W, rh, rw, tiles_y, tiles_x);
    else
        upsample2x_fwd_kernel<false><<<(unsigned)blocks, UF_NT, 0, st>>>(in, out, H, W, rh, rw, tiles_y, tiles_x);
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
    const int tiles_y = ceil_div(H, UB_TH), tiles_x = ceil_div(W, UB_TW);
    const long long blocks = N * tiles_y * tiles_x;
    TAI_REQUIRE(blocks < (1LL << 31), TAI_ERR_TOO_LARGE, "upsample_bilinear2x_backward_b200: too many tiles");
    TimingScope ts("upsample2x_bwd", st, 0.0, 4.0 * (5 * in_el));  // read the 2H x 2W gradient once, write H x W
    upsample2x_bwd_kernel<<<(unsigned)blocks, UB_NT, 0, st>>>(grad_out, grad_in, H, W, rh, rw, tiles_y, tiles_x);
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
