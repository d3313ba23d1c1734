// Bias + activation epilogue of the convolution layers for sm_100a, and its adjoint.
//
// Every convolution of the TAI / MC-Net / SloMo stacks is `Conv2d(bias=True)` followed by ReLU / LeakyReLU (or
// nothing): mcnet.py:28-45,79-104,137-225; tai.py:244-347; slomo.py:28-260.  The library evaluates such a layer as
// FOUR streaming passes around the cuDNN kernel -- forward: a broadcast `add_(bias)` (torch's non-vectorised
// elementwise kernel: 194 us for [64,64,128,128]) and `clamp_min`; backward: `threshold_backward` and a
// `sum` over (N,H,W) for the bias gradient -- 785 + 1716 launches and ~60 ms of the 655 ms KTH training step
// (profiles/r01_step_kernels.csv).  Here: ONE pass each way.
//
//   forward  (in place on the convolution's output y [N,C,HW]):  y = act(y + b[c])
//   backward:  gin = gout * act'(y_out)   and   gb[c] = sum_{n,hw} gin        (for act = none gin == gout: nothing is
//              written and only the reduction runs)
//
// act' is taken from the OUTPUT (relu: out > 0; leaky: out > 0 ? 1 : alpha -- same sign as the pre-activation for
// alpha > 0), which is what the library's threshold_backward / leaky_relu_backward do with `result`.
// The forward is the same two FP32 operations the library performs (one add, one max / select): bit-identical.
// The bias gradient is a two-stage fixed-order sum (per-CTA partials, then one thread block per channel in double):
// deterministic.
#include "common.cuh"

namespace tai {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2 };

__device__ __forceinline__ float act_fwd(float v, int act, float alpha)
{
    if (act == ACT_RELU) return v < 0.f ? 0.f : v;               // clamp_min(0); NaN propagates like the library's
    if (act == ACT_LEAKY) return v > 0.f ? v : v * alpha;
    return v;
}

__device__ __forceinline__ float act_grad(float g, float out, int act, float alpha)
{
    if (act == ACT_RELU) return out <= 0.f ? 0.f : g;            // threshold_backward(grad, result, 0)
    if (act == ACT_LEAKY) return out > 0.f ? g : g * alpha;      // leaky_relu_backward on the result
    return g;
}

// one row = HW contiguous elements of one (n, c) plane; a thread block walks rows with float4 accesses
template <bool VEC>
__global__ void __launch_bounds__(256)
bias_act_fwd_kernel(float *__restrict__ y, const float *__restrict__ bias, long rows, int C, int HW, int act, float alpha)
{
    const int per = VEC ? HW / 4 : HW;
    const long total = rows * per;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long row = idx / per;
        const float b = __ldg(bias + (int)(row % C));
        if (VEC) {
            float4 v = reinterpret_cast<float4 *>(y)[idx];
            v.x = act_fwd(v.x + b, act, alpha);
            v.y = act_fwd(v.y + b, act, alpha);
            v.z = act_fwd(v.z + b, act, alpha);
            v.w = act_fwd(v.w + b, act, alpha);
            reinterpret_cast<float4 *>(y)[idx] = v;
        } else {
            y[idx] = act_fwd(y[idx] + b, act, alpha);
        }
    }
}

// grid (chunks, C): CTA (k, c) owns the planes n = k, k + chunks, ... of channel c; writes gin (may alias gout) and
// its partial bias-gradient sum to partial[c * chunks + k]
template <bool VEC>
__global__ void __launch_bounds__(256)
bias_act_bwd_kernel(const float *gout, const float *__restrict__ out, float *gin, float *__restrict__ partial, int N, int C,
                    int HW, int act, float alpha, int write_gin)
{
    __shared__ float s_red[8];
    const int c = blockIdx.y, chunks = gridDim.x;
    float acc = 0.f;
    if (VEC) {
        // The CTA's planes n = k, k + chunks, ... form one flat list of 128-bit items (q per plane); four items,
        // 1024 apart, are in flight per thread whatever the plane size (a 32 x 32 plane is a single item per
        // thread: walking plane by plane left one load pair in flight and ran at 61 % of the copy bandwidth).
        const int q = HW / 4;
        const int nplanes = (N - (int)blockIdx.x + chunks - 1) / chunks;
        const int items = nplanes * q;   // < 2^29: the whole tensor has < 2^31 elements
        auto addr = [&](int i) -> long {
            const int pl = i / q;
            return ((long)(blockIdx.x + pl * chunks) * C + c) * q + (i - pl * q);
        };
        const float4 *g4 = reinterpret_cast<const float4 *>(gout);
        const float4 *o4 = reinterpret_cast<const float4 *>(out);
        float4 *d4 = reinterpret_cast<float4 *>(gin);
        int i = threadIdx.x;
        for (; i + 768 < items; i += 1024) {
            long a[4];
            float4 g[4], o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a[u] = addr(i + 256 * u);
                g[u] = g4[a[u]];
            }
            if (act != ACT_NONE) {
#pragma unroll
                for (int u = 0; u < 4; ++u) o[u] = o4[a[u]];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    g[u].x = act_grad(g[u].x, o[u].x, act, alpha);
                    g[u].y = act_grad(g[u].y, o[u].y, act, alpha);
                    g[u].z = act_grad(g[u].z, o[u].z, act, alpha);
                    g[u].w = act_grad(g[u].w, o[u].w, act, alpha);
                    if (write_gin) d4[a[u]] = g[u];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += (g[u].x + g[u].y) + (g[u].z + g[u].w);
        }
        for (; i < items; i += 256) {
            const long a = addr(i);
            float4 g = g4[a];
            if (act != ACT_NONE) {
                const float4 o = o4[a];
                g.x = act_grad(g.x, o.x, act, alpha);
                g.y = act_grad(g.y, o.y, act, alpha);
                g.z = act_grad(g.z, o.z, act, alpha);
                g.w = act_grad(g.w, o.w, act, alpha);
                if (write_gin) d4[a] = g;
            }
            acc += (g.x + g.y) + (g.z + g.w);
        }
    } else {
        for (int n = blockIdx.x; n < N; n += chunks) {
            const long base = ((long)n * C + c) * HW;
            for (int i = threadIdx.x; i < HW; i += 256) {
                float g = gout[base + i];
                if (act != ACT_NONE) {
                    g = act_grad(g, out[base + i], act, alpha);
                    if (write_gin) gin[base + i] = g;
                }
                acc += g;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_red[w];
        partial[(long)c * chunks + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(32)
bias_grad_finalize_kernel(const float *__restrict__ partial, float *__restrict__ gbias, int C, int chunks)
{
    const int c = blockIdx.x * 32 + threadIdx.x;
    if (c >= C) return;
    double t = 0.0;
    for (int k = 0; k < chunks; ++k) t += (double)partial[(long)c * chunks + k];
    gbias[c] = (float)t;
}

static inline int bias_chunks(int N, int C)
{
    // enough CTAs to fill the chip (~8 per SM), at most one per plane -- and a divisor of N where one is close, so
    // that every CTA of a channel owns the same number of planes (19 chunks over 64 planes: 4 vs 3.4 on average)
    int chunks = (sm_count() * 8 + C - 1) / C;
    if (chunks > N) chunks = N;
    if (chunks < 1) chunks = 1;
    for (int d = chunks; d >= (chunks + 1) / 2; --d)
        if (N % d == 0) return d;
    return chunks;
}

// v / (sqrt(sum v^2) + eps) for one short vector: the `_l2normalize` of the spectral-norm power iteration
// (SNDiscriminator.py:5-7: pow, sum, pow, add, div = five launches, 1170 times per KTH training step).  One CTA.
__global__ void __launch_bounds__(256)
l2_normalize_kernel(const float *__restrict__ v, float *__restrict__ out, int n, float eps)
{
    __shared__ float s_red[8];
    __shared__ float s_inv;
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) {
        const float x = v[i];
        acc = fmaf(x, x, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_red[w];
        s_inv = sqrtf(t) + eps;
    }
    __syncthreads();
    const float d = s_inv;
    for (int i = threadIdx.x; i < n; i += 256) out[i] = v[i] / d;
}

}  // namespace tai

using namespace tai;

extern "C" int bias_act_forward_b200(float *y, const float *bias, long long N, int C, int HW, int act, float alpha, void *stream)
{
    TAI_REQUIRE(y && bias && N > 0 && C > 0 && HW > 0 && act >= ACT_NONE && act <= ACT_LEAKY, TAI_ERR_INVALID_ARGUMENT,
                "bias_act_forward_b200: bad arguments N=%lld C=%d HW=%d act=%d", N, C, HW, act);
    TAI_REQUIRE(fits_int31(N * (long long)C * HW), TAI_ERR_TOO_LARGE, "bias_act_forward_b200: tensor has >= 2^31 elements");
    cudaStream_t st = (cudaStream_t)stream;
    const long rows = (long)N * C;
    const bool vec = (HW % 4 == 0) && (((uintptr_t)y & 15) == 0);
    const long items = rows * (vec ? HW / 4 : HW);
    long grid = (items + 255) / 256;
    const long cap = (long)sm_count() * 16;
    if (grid > cap) grid = cap;
    TimingScope ts("bias_act_fwd", st, 0.0, 8.0 * rows * HW);  // read + write the activation once
    if (vec)
        bias_act_fwd_kernel<true><<<(unsigned)grid, 256, 0, st>>>(y, bias, rows, C, HW, act, alpha);
    else
        bias_act_fwd_kernel<false><<<(unsigned)grid, 256, 0, st>>>(y, bias, rows, C, HW, act, alpha);
    return check_launch("bias_act_fwd_kernel");
}

extern "C" long long bias_act_backward_workspace_bytes(long long N, int C)
{
    if (N <= 0 || C <= 0) return 0;
    const int n = N > (1 << 30) ? (1 << 30) : (int)N;
    return 4LL * C * bias_chunks(n, C);
}

extern "C" int bias_act_backward_b200(const float *grad_out, const float *out, float *grad_in, float *grad_bias, void *workspace,
                                      long long N, int C, int HW, int act, float alpha, void *stream)
{
    TAI_REQUIRE(grad_out && grad_bias && workspace && N > 0 && N < (1LL << 30) && C > 0 && C <= 65535 && HW > 0 &&
                    act >= ACT_NONE && act <= ACT_LEAKY,
                TAI_ERR_INVALID_ARGUMENT, "bias_act_backward_b200: bad arguments N=%lld C=%d HW=%d act=%d", N, C, HW, act);
    TAI_REQUIRE(act == ACT_NONE || out != nullptr, TAI_ERR_INVALID_ARGUMENT, "bias_act_backward_b200: the activation needs the forward output");
    TAI_REQUIRE(fits_int31(N * (long long)C * HW), TAI_ERR_TOO_LARGE, "bias_act_backward_b200: tensor has >= 2^31 elements");
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = bias_chunks((int)N, C);
    const int write_gin = (act != ACT_NONE && grad_in != nullptr) ? 1 : 0;
    float *gin = grad_in ? grad_in : const_cast<float *>(grad_out);
    const bool vec = (HW % 4 == 0) && (((uintptr_t)grad_out | (uintptr_t)gin | (uintptr_t)out) & 15) == 0;
    float *partial = reinterpret_cast<float *>(workspace);
    const double el = (double)N * C * HW;
    TimingScope ts("bias_act_bwd", st, 0.0, 4.0 * el * (act == ACT_NONE ? 1.0 : 3.0));  // both launches
    const dim3 grid((unsigned)chunks, (unsigned)C);
    if (vec)
        bias_act_bwd_kernel<true><<<grid, 256, 0, st>>>(grad_out, out, gin, partial, (int)N, C, HW, act, alpha, write_gin);
    else
        bias_act_bwd_kernel<false><<<grid, 256, 0, st>>>(grad_out, out, gin, partial, (int)N, C, HW, act, alpha, write_gin);
    int rc = check_launch("bias_act_bwd_kernel");
    if (rc != TAI_OK) return rc;
    bias_grad_finalize_kernel<<<(C + 31) / 32, 32, 0, st>>>(partial, grad_bias, C, chunks);
    return check_launch("bias_grad_finalize_kernel");
}

extern "C" int l2_normalize_b200(const float *v, float *out, int n, float eps, void *stream)
{
    TAI_REQUIRE(v && out && n > 0, TAI_ERR_INVALID_ARGUMENT, "l2_normalize_b200: bad arguments");
    // no TimingScope: 1014 launches of ~3 us per training step -- the two CUDA events per launch cost more than the kernel
    l2_normalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(v, out, n, eps);
    return check_launch("l2_normalize_kernel");
}
