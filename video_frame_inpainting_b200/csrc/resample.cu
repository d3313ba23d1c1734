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

// out[n, oy, ox] = w0y * (w0x * I[y0,x0] + w1x * I[y0,x1]) + w1y * (w0x * I[y1,x0] + w1x * I[y1,x1])
// A thread produces a block of UP_R output rows x VEC output columns: the column taps are formed once and
// reused for every row (the tap arithmetic, not the memory system, bounded the one-element-per-thread form).
constexpr int UP_R = 4;
template <int VEC>
__global__ void __launch_bounds__(256)
upsample2x_fwd_kernel(const float *__restrict__ in, float *__restrict__ out, long N, int H, int W, float rh, float rw)
{
    const int Ho = 2 * H, Wo = 2 * W;
    const int wq = Wo / VEC;
    const int hq = (Ho + UP_R - 1) / UP_R;
    const long total = N * hq * wq;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int q = (int)(idx % wq);
        const long t = idx / wq;
        const int oy0 = (int)(t % hq) * UP_R;
        const long n = t / hq;
        Tap2 tx[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) tx[k] = up_tap(q * VEC + k, rw, W);
        const float *plane = in + n * H * W;
#pragma unroll
        for (int r = 0; r < UP_R; ++r) {
            const int oy = oy0 + r;
            if (oy < Ho) {
                const Tap2 ty = up_tap(oy, rh, H);
                const float *r0 = plane + (long)ty.i0 * W;
                const float *r1 = plane + (long)ty.i1 * W;
                float res[VEC];
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const float top = tx[k].w0 * __ldg(r0 + tx[k].i0) + tx[k].w1 * __ldg(r0 + tx[k].i1);
                    const float bot = tx[k].w0 * __ldg(r1 + tx[k].i0) + tx[k].w1 * __ldg(r1 + tx[k].i1);
                    res[k] = ty.w0 * top + ty.w1 * bot;
                }
                float *o = out + (n * Ho + oy) * Wo + (long)q * VEC;
                if (VEC == 4)
                    *reinterpret_cast<float4 *>(o) = make_float4(res[0], res[1], res[2], res[3]);
                else
                    o[0] = res[0];
            }
        }
    }
}

// Adjoint as a gather (deterministic, no atomics): input pixel (y, x) collects every output pixel whose
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

// A thread owns one input column and UB_R consecutive input rows: the 6 column weights are formed once, the
// 2*UB_R+4 candidate output rows are reduced along x once (row sums) and shared by the UB_R input rows.
constexpr int UB_R = 4;
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const float *__restrict__ gout, float *__restrict__ gin, long N, int H, int W, float rh, float rw)
{
    const int Ho = 2 * H, Wo = 2 * W;
    const int hq = (H + UB_R - 1) / UB_R;
    const long total = N * hq * W;
    constexpr int NC = 2 * UB_R + 4;  // candidate output rows 2*y0-2 .. 2*y0+2*UB_R+1
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % W);
        const long t = idx / W;
        const int y0 = (int)(t % hq) * UB_R;
        const long n = t / hq;
        const int ox0 = 2 * x - 2;
        float wx[6];
#pragma unroll
        for (int b = 0; b < 6; ++b) wx[b] = up_adjoint_weight(ox0 + b, x, rw, W, Wo);
        const float *g = gout + n * Ho * Wo;
        const int oyb = 2 * y0 - 2;
        float rs[NC];
#pragma unroll
        for (int a = 0; a < NC; ++a) {
            const int oy = oyb + a;
            float racc = 0.f;
            if (oy >= 0 && oy < Ho) {
                const float *grow = g + (long)oy * Wo + ox0;
#pragma unroll
                for (int b = 0; b < 6; ++b)
                    if (wx[b] != 0.f) racc = fmaf(wx[b], __ldg(grow + b), racc);
            }
            rs[a] = racc;
        }
#pragma unroll
        for (int r = 0; r < UB_R; ++r) {
            const int y = y0 + r;
            if (y < H) {
                float acc = 0.f;
#pragma unroll
                for (int a = 0; a < 6; ++a)  // output rows 2y-2 .. 2y+3 = candidates 2r .. 2r+5
                    acc = fmaf(up_adjoint_weight(2 * y - 2 + a, y, rh, H, Ho), rs[2 * r + a], acc);
                gin[(n * H + y) * W + x] = acc;
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
    TimingScope ts("upsample2x_fwd", st, 0.0, 4.0 * (out_el + out_el / 4));  // read the input once, write the result
    if ((W % 2) == 0 && aligned16(out, out))
        upsample2x_fwd_kernel<4><<<resample_grid((long)N * ((2 * H + UP_R - 1) / UP_R) * (2 * W / 4), 256), 256, 0, st>>>(in, out, (long)N, H, W, rh, rw);
    else
        upsample2x_fwd_kernel<1><<<resample_grid((long)N * ((2 * H + UP_R - 1) / UP_R) * (2 * W), 256), 256, 0, st>>>(in, out, (long)N, H, W, rh, rw);
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
    TimingScope ts("upsample2x_bwd", st, 0.0, 4.0 * (5 * in_el));
    upsample2x_bwd_kernel<<<resample_grid((long)N * ((H + UB_R - 1) / UB_R) * W, 256), 256, 0, st>>>(grad_out, grad_in, (long)N, H, W, rh, rw);
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
