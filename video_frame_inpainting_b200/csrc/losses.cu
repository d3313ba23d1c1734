// Reconstruction losses of the TAI training step for sm_100a:
// MSELoss + GDL of a prediction against the ground truth, with the [-1,1] -> [0,1] inverse transform folded in.
//
//   reference: environments.py:363-371, 447-451 (three prediction tensors per step, each: permute +
//   contiguous + inverse_transform twice, MSELoss, GDL) and losses.py:24-45 (four sliced differences, two
//   L1Loss(reduce=False), two sliced contiguous copies, add, mean): ~25 elementwise / copy / reduction launches
//   per prediction tensor forward and about as many backward, each streaming the 10 MB tensors again.
//   Here: one streaming kernel + a 1-CTA finalize forward (8 B per element), one streaming kernel backward
//   (12 B per element).  Both are pure HBM streams.
//
// Arithmetic follows the reference's operation order in FP32 so that every |.| term, and in particular every
// sign() the backward selects, is the one the reference computes:
//   v01 = (v + add) * mul                     util.py:22-23 ((images + 1.) / 2; * 0.5 is exact)
//   mse term   (x01 - y01)^2                  mean over planes*H*W
//   w term     |(x01[r,c] - x01[r,c+1]) - (y01[r,c] - y01[r,c+1])|   r in 1..H-1, c in 0..W-2   (losses.py:30,32,34)
//   h term     |(x01[r+1,c] - x01[r,c]) - (y01[r+1,c] - y01[r,c])|   r in 0..H-2, c in 1..W-1   (losses.py:31,33,35)
//   gdl = mean over planes*(H-1)*(W-1) of (w + h)                                                (losses.py:41-43)
// The means do not depend on the order of the planes, so the reference's time-major permute is not needed.
#include "common.cuh"

namespace tai {

constexpr int LS_NT = 256;

__device__ __forceinline__ float block_sum(float v, float *s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) s_red[w] = v;
    __syncthreads();
    float t = (threadIdx.x < LS_NT / 32) ? s_red[threadIdx.x] : 0.f;
    if (w == 0) {
#pragma unroll
        for (int o = LS_NT / 64; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;  // valid in thread 0
}

template <int VEC>
struct RowLoad {
    float x[VEC], y[VEC];
};

template <int VEC>
__device__ __forceinline__ void load_row(const float *__restrict__ px, const float *__restrict__ py, float add, float mul,
                                         RowLoad<VEC> &o)
{
    if (VEC == 4) {
        const float4 a = ld_stream4(reinterpret_cast<const float4 *>(px));
        const float4 b = ld_stream4(reinterpret_cast<const float4 *>(py));
        o.x[0] = a.x; o.x[1] = a.y; o.x[2] = a.z; o.x[3] = a.w;
        o.y[0] = b.x; o.y[1] = b.y; o.y[2] = b.z; o.y[3] = b.w;
    } else {
        o.x[0] = __ldg(px);
        o.y[0] = __ldg(py);
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        o.x[k] = __fmul_rn(__fadd_rn(o.x[k], add), mul);
        o.y[k] = __fmul_rn(__fadd_rn(o.y[k], add), mul);
    }
}

__device__ __forceinline__ float tf(const float *p, float add, float mul) { return __fmul_rn(__fadd_rn(__ldg(p), add), mul); }

// |(a - b) - (c - d)| with one rounding per reference operation
__device__ __forceinline__ float gdiff(float a, float b, float c, float d)
{
    return __fsub_rn(__fsub_rn(a, b), __fsub_rn(c, d));
}

// A thread owns VEC consecutive columns of one plane and walks the rows [r0, r1) downwards; the row below the
// current one is loaded one step ahead (it is needed for the h term and becomes the current row).
template <int VEC>
__global__ void __launch_bounds__(LS_NT)
l2_gdl_fwd_kernel(const float *__restrict__ x, const float *__restrict__ y, long planes, int H, int W, float add, float mul,
                  int R, float2 *__restrict__ partials)
{
    __shared__ float s_red[LS_NT / 32];
    const int wq = W / VEC;
    const long item = (long)blockIdx.x * LS_NT + threadIdx.x;
    float sq = 0.f, gd = 0.f;
    if (item < planes * wq) {
        const long n = item / wq;
        const int c0 = (int)(item - n * wq) * VEC;
        const int r0 = blockIdx.y * R, r1 = min(H, r0 + R);
        const float *px = x + (n * H + r0) * W + c0;
        const float *py = y + (n * H + r0) * W + c0;
        const bool has_right = c0 + VEC < W;
        RowLoad<VEC> cur, nxt;
        load_row<VEC>(px, py, add, mul, cur);
        for (int r = r0; r < r1; ++r) {
            const bool below = r + 1 < H;
            if (below) load_row<VEC>(px + W, py + W, add, mul, nxt);
            float xr = 0.f, yr = 0.f;
            if (has_right && r >= 1) {
                xr = tf(px + VEC, add, mul);
                yr = tf(py + VEC, add, mul);
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const float d = __fsub_rn(cur.x[k], cur.y[k]);
                sq = fmaf(d, d, sq);
                if (r >= 1 && c0 + k <= W - 2) {
                    const float xn = (k + 1 < VEC) ? cur.x[k + 1 < VEC ? k + 1 : 0] : xr;
                    const float yn = (k + 1 < VEC) ? cur.y[k + 1 < VEC ? k + 1 : 0] : yr;
                    gd += fabsf(gdiff(cur.x[k], xn, cur.y[k], yn));
                }
                if (below && c0 + k >= 1) gd += fabsf(gdiff(nxt.x[k], cur.x[k], nxt.y[k], cur.y[k]));
            }
            cur = nxt;
            px += W;
            py += W;
        }
    }
    const float a = block_sum(sq, s_red);
    const float b = block_sum(gd, s_red);
    if (threadIdx.x == 0) partials[(long)blockIdx.y * gridDim.x + blockIdx.x] = make_float2(a, b);
}

// out[0] = sum(sq) / n_mse, out[1] = sum(gd) / n_gdl; fixed summation order (deterministic), in double
__global__ void __launch_bounds__(LS_NT)
l2_gdl_finalize_kernel(const float2 *__restrict__ partials, long count, double inv_mse, double inv_gdl, float *__restrict__ out)
{
    __shared__ double s_a[LS_NT], s_b[LS_NT];
    double a = 0.0, b = 0.0;
    for (long i = threadIdx.x; i < count; i += LS_NT) {
        const float2 p = partials[i];
        a += (double)p.x;
        b += (double)p.y;
    }
    s_a[threadIdx.x] = a;
    s_b[threadIdx.x] = b;
    __syncthreads();
    for (int o = LS_NT / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s_a[threadIdx.x] += s_a[threadIdx.x + o];
            s_b[threadIdx.x] += s_b[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = (float)(s_a[0] * inv_mse);
        out[1] = (float)(s_b[0] * inv_gdl);  // 0 * inf = NaN for H == 1 or W == 1: the mean of an empty tensor
    }
}

__device__ __forceinline__ float sgn(float v) { return (v > 0.f ? 1.f : 0.f) - (v < 0.f ? 1.f : 0.f); }

// d loss / d x with  loss = g_mse * mse + g_gdl * gdl  (g_* are device scalars: the upstream gradients of
// the two returned means, read on the device so that no host synchronisation is needed).
//   grad[r,c] = cm * (x01 - y01)
//             + cg * (  [r>=1, c<=W-2] sW(r,c) - [r>=1, c>=1] sW(r,c-1) + [c>=1, r>=1] sH(r-1,c) - [c>=1, r<=H-2] sH(r,c) )
//   cm = g_mse * 2 * mul / n_mse,  cg = g_gdl * mul / n_gdl,  sW / sH = sign of the w / h term's argument.
template <int VEC>
__global__ void __launch_bounds__(LS_NT)
l2_gdl_bwd_kernel(const float *__restrict__ x, const float *__restrict__ y, long planes, int H, int W, float add, float mul,
                  int R, const float *__restrict__ g_mse, const float *__restrict__ g_gdl, float k_mse, float k_gdl,
                  float *__restrict__ gx)
{
    const int wq = W / VEC;
    const long item = (long)blockIdx.x * LS_NT + threadIdx.x;
    if (item >= planes * wq) return;
    const float cm = (g_mse ? __ldg(g_mse) : 0.f) * k_mse;
    const float cg = (g_gdl ? __ldg(g_gdl) : 0.f) * k_gdl;
    const long n = item / wq;
    const int c0 = (int)(item - n * wq) * VEC;
    const int r0 = blockIdx.y * R, r1 = min(H, r0 + R);
    const float *px = x + (n * H + r0) * W + c0;
    const float *py = y + (n * H + r0) * W + c0;
    float *pg = gx + (n * H + r0) * W + c0;
    const bool has_left = c0 >= 1, has_right = c0 + VEC < W;
    RowLoad<VEC> prv, cur, nxt;
    load_row<VEC>(px, py, add, mul, cur);
    prv = cur;
    if (r0 >= 1) load_row<VEC>(px - W, py - W, add, mul, prv);
    for (int r = r0; r < r1; ++r) {
        const bool below = r + 1 < H;
        nxt = cur;
        if (below) load_row<VEC>(px + W, py + W, add, mul, nxt);
        float xl = 0.f, yl = 0.f, xr = 0.f, yr = 0.f;
        if (r >= 1) {
            if (has_left) {
                xl = tf(px - 1, add, mul);
                yl = tf(py - 1, add, mul);
            }
            if (has_right) {
                xr = tf(px + VEC, add, mul);
                yr = tf(py + VEC, add, mul);
            }
        }
        // sw[k] = sW(r, c0 + k - 1), k = 0 .. VEC: the pairs (c0-1,c0) .. (c0+VEC-1, c0+VEC); zero where the pair does not exist
        float ex[VEC + 2], ey[VEC + 2];
        ex[0] = xl;
        ey[0] = yl;
        ex[VEC + 1] = xr;
        ey[VEC + 1] = yr;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            ex[k + 1] = cur.x[k];
            ey[k + 1] = cur.y[k];
        }
        float sw[VEC + 1];
#pragma unroll
        for (int k = 0; k <= VEC; ++k) {
            const int c = c0 + k - 1;  // left column of the pair
            sw[k] = (r >= 1 && c >= 0 && c <= W - 2) ? sgn(gdiff(ex[k], ex[k + 1], ey[k], ey[k + 1])) : 0.f;
        }
        float g[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const int c = c0 + k;
            float s = sw[k + 1] - sw[k];
            if (c >= 1) {
                if (r >= 1) s += sgn(gdiff(cur.x[k], prv.x[k], cur.y[k], prv.y[k]));
                if (below) s -= sgn(gdiff(nxt.x[k], cur.x[k], nxt.y[k], cur.y[k]));
            }
            g[k] = fmaf(cg, s, cm * __fsub_rn(cur.x[k], cur.y[k]));
        }
        if (VEC == 4)
            *reinterpret_cast<float4 *>(pg) = make_float4(g[0], g[1], g[2], g[3]);
        else
            pg[0] = g[0];
        prv = cur;
        cur = nxt;
        px += W;
        py += W;
        pg += W;
    }
}

static inline int loss_rows_per_thread(long planes, int wq, int H)
{
    int R = 32;  // rows per thread; shorter walks when the tensor would not fill the chip
    while (R > 4 && planes * wq * ceil_div(H, R) < (long)sm_count() * 2048) R /= 2;
    return R;
}

static inline bool loss_vec4(const void *a, const void *b, const void *c, int W)
{
    return (W % 4) == 0 && ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0);
}

static int loss_args_ok(const char *who, const void *a, const void *b, const void *c, long long planes, int H, int W)
{
    TAI_REQUIRE(a && b && c && planes > 0 && H > 0 && W > 0, TAI_ERR_INVALID_ARGUMENT, "%s: bad arguments planes=%lld H=%d W=%d", who,
                planes, H, W);
    TAI_REQUIRE(fits_int31(planes * (long long)H * W), TAI_ERR_TOO_LARGE, "%s: tensor has >= 2^31 elements", who);
    return TAI_OK;
}

}  // namespace tai

using namespace tai;

extern "C" long long l2_gdl_loss_workspace_bytes(long long planes, int H, int W)
{
    if (planes <= 0 || H <= 0 || W <= 0) return 0;
    // upper bound over both vector widths and every rows-per-thread choice: one float2 per CTA
    const long long bx = (planes * W + LS_NT - 1) / LS_NT;
    return 8 * bx * ceil_div(H, 4);
}

extern "C" int l2_gdl_loss_forward_b200(const float *pred, const float *target, long long planes, int H, int W, float add,
                                        float mul, float *out2, void *workspace, void *stream)
{
    int rc = loss_args_ok("l2_gdl_loss_forward_b200", pred, target, out2, planes, H, W);
    if (rc != TAI_OK) return rc;
    TAI_REQUIRE(workspace != nullptr, TAI_ERR_INVALID_ARGUMENT, "l2_gdl_loss_forward_b200: null workspace");
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = loss_vec4(pred, target, nullptr, W);
    const int wq = v4 ? W / 4 : W;
    const int R = loss_rows_per_thread((long)planes, wq, H);
    const long long bx = (planes * wq + LS_NT - 1) / LS_NT;
    const int by = ceil_div(H, R);
    TAI_REQUIRE(bx < (1LL << 31) && by < 65536, TAI_ERR_TOO_LARGE, "l2_gdl_loss_forward_b200: grid too large");
    const double n_mse = (double)planes * H * W, n_gdl = (double)planes * (H - 1) * (W - 1);
    float2 *partials = reinterpret_cast<float2 *>(workspace);
    TimingScope ts("l2_gdl_fwd", st, 0.0, 8.0 * n_mse);  // read the prediction and the target once; both launches
    if (v4)
        l2_gdl_fwd_kernel<4><<<dim3((unsigned)bx, (unsigned)by), LS_NT, 0, st>>>(pred, target, (long)planes, H, W, add, mul, R, partials);
    else
        l2_gdl_fwd_kernel<1><<<dim3((unsigned)bx, (unsigned)by), LS_NT, 0, st>>>(pred, target, (long)planes, H, W, add, mul, R, partials);
    rc = check_launch("l2_gdl_fwd_kernel");
    if (rc != TAI_OK) return rc;
    l2_gdl_finalize_kernel<<<1, LS_NT, 0, st>>>(partials, (long)(bx * by), 1.0 / n_mse, 1.0 / n_gdl, out2);
    return check_launch("l2_gdl_finalize_kernel");
}

extern "C" int l2_gdl_loss_backward_b200(const float *pred, const float *target, long long planes, int H, int W, float add,
                                         float mul, const float *grad_mse, const float *grad_gdl, float *grad_pred,
                                         void *stream)
{
    int rc = loss_args_ok("l2_gdl_loss_backward_b200", pred, target, grad_pred, planes, H, W);
    if (rc != TAI_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = loss_vec4(pred, target, grad_pred, W);
    const int wq = v4 ? W / 4 : W;
    const int R = loss_rows_per_thread((long)planes, wq, H);
    const long long bx = (planes * wq + LS_NT - 1) / LS_NT;
    const int by = ceil_div(H, R);
    TAI_REQUIRE(bx < (1LL << 31) && by < 65536, TAI_ERR_TOO_LARGE, "l2_gdl_loss_backward_b200: grid too large");
    const double n_mse = (double)planes * H * W, n_gdl = (double)planes * (H - 1) * (W - 1);
    const float k_mse = (float)(2.0 * (double)mul / n_mse);
    const float k_gdl = n_gdl > 0 ? (float)((double)mul / n_gdl) : 0.f;
    TimingScope ts("l2_gdl_bwd", st, 0.0, 12.0 * n_mse);  // read the prediction and the target, write the gradient
    if (v4)
        l2_gdl_bwd_kernel<4><<<dim3((unsigned)bx, (unsigned)by), LS_NT, 0, st>>>(pred, target, (long)planes, H, W, add, mul, R,
                                                                                 grad_mse, grad_gdl, k_mse, k_gdl, grad_pred);
    else
        l2_gdl_bwd_kernel<1><<<dim3((unsigned)bx, (unsigned)by), LS_NT, 0, st>>>(pred, target, (long)planes, H, W, add, mul, R,
                                                                                 grad_mse, grad_gdl, k_mse, k_gdl, grad_pred);
    return check_launch("l2_gdl_bwd_kernel");
}
