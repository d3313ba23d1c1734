// HBM-bound kernels of the TAI / Super-SloMo path for sm_100a: ConvLSTM gates, bilinear backward
// warp, SloMo flow-combine / refine-blend fusions, and the FFMA probe used by bench.py.
// Every kernel streams its operands exactly once with 128-bit accesses where alignment allows.
#include "common.cuh"
#include "warp.cuh"

namespace tai {


// ------------------------------------------------------------------------------------------------
// ConvLSTM gates (mcnet.py:287-293).  conv_out [B,4F,HW] holds (i,j,f,o) as four contiguous F*HW
// slabs per sample; state [B,2F,HW] holds (c,h).  28 B of traffic per state element.
//
// Activations: the accurate expf / tanhf / IEEE-division sequences cost ~170 instructions per state element,
// which made the forward kernel ISSUE-bound (63 % of the copy bandwidth, same time as the backward kernel that
// moves twice the bytes).  ex2.approx / rcp.approx based forms are ~3x shorter; their error (a few ulp of the
// exponential, <= 2e-6 relative for |x| < 20) is far inside the 1e-4 parity bar.  tanh needs care near zero,
// where (1 - t) / (1 + t) cancels: below 0.25 an odd polynomial (next term 9e-3 * x^11) is selected.
__device__ __forceinline__ float ex2_ftz(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_ftz(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 1 / (1 + e^-x): e^-x = +inf for x < -88.7 -> rcp(inf) = 0, the limit
__device__ __forceinline__ float sigmoid_acc(float x) { return rcp_ftz(1.f + ex2_ftz(-1.4426950408889634f * x)); }

__device__ __forceinline__ float tanh_acc(float x)
{
    const float ax = fabsf(x);
    const float t = ex2_ftz(-2.8853900817779268f * ax);
    const float big = (1.f - t) * rcp_ftz(1.f + t);
    const float x2 = ax * ax;
    float p = fmaf(x2, 62.f / 2835.f, -17.f / 315.f);
    p = fmaf(x2, p, 2.f / 15.f);
    p = fmaf(x2, p, -1.f / 3.f);
    p = fmaf(x2 * ax, p, ax);
    return copysignf(ax < 0.25f ? p : big, x);
}

template <int VEC>
struct GateItem {
    float gi[VEC], gj[VEC], gf[VEC], go[VEC], c[VEC];
    float *ob;
};

template <int VEC>
__device__ __forceinline__ void gates_load(GateItem<VEC> &it, const float *__restrict__ conv, const float *__restrict__ state,
                                           float *__restrict__ nstate, long idx, long per, long slab)
{
    const long b = idx / per;
    const long e = (idx - b * per) * VEC;
    const float *cb = conv + b * 4 * slab + e;
    const float *sb = state + b * 2 * slab + e;
    it.ob = nstate + b * 2 * slab + e;
    if (VEC == 4) {
        *reinterpret_cast<float4 *>(it.gi) = ld_stream4(reinterpret_cast<const float4 *>(cb));
        *reinterpret_cast<float4 *>(it.gj) = ld_stream4(reinterpret_cast<const float4 *>(cb + slab));
        *reinterpret_cast<float4 *>(it.gf) = ld_stream4(reinterpret_cast<const float4 *>(cb + 2 * slab));
        *reinterpret_cast<float4 *>(it.go) = ld_stream4(reinterpret_cast<const float4 *>(cb + 3 * slab));
        *reinterpret_cast<float4 *>(it.c) = ld_stream4(reinterpret_cast<const float4 *>(sb));
    } else {
        it.gi[0] = cb[0]; it.gj[0] = cb[slab]; it.gf[0] = cb[2 * slab]; it.go[0] = cb[3 * slab]; it.c[0] = sb[0];
    }
}

template <int VEC>
__device__ __forceinline__ void gates_finish(const GateItem<VEC> &it, long slab, float fb)
{
    float nc[VEC], nh[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        nc[k] = it.c[k] * sigmoid_acc(it.gf[k] + fb) + sigmoid_acc(it.gi[k]) * tanh_acc(it.gj[k]);
        nh[k] = tanh_acc(nc[k]) * sigmoid_acc(it.go[k]);
    }
    if (VEC == 4) {
        *reinterpret_cast<float4 *>(it.ob) = *reinterpret_cast<float4 *>(nc);
        *reinterpret_cast<float4 *>(it.ob + slab) = *reinterpret_cast<float4 *>(nh);
    } else {
        it.ob[0] = nc[0];
        it.ob[slab] = nh[0];
    }
}

// Two grid-stride trips are in flight per thread: the ten 128-bit loads of both items are issued before the
// first activation is evaluated (ncu: the one-item form sat at 40 % of the DRAM bandwidth with
// long_scoreboard = 10.7 stall cycles per issue -- not enough bytes in flight for a 10 us kernel).
template <int VEC>
__global__ void __launch_bounds__(128, 8)
gates_fwd_kernel(const float *__restrict__ conv, const float *__restrict__ state, float *__restrict__ nstate,
                 int B, long slab, float fb)
{
    const long per = slab / VEC;
    const long n = (long)B * per;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += 2 * stride) {
        GateItem<VEC> a, b;
        const bool two = idx + stride < n;
        gates_load<VEC>(a, conv, state, nstate, idx, per, slab);
        if (two) gates_load<VEC>(b, conv, state, nstate, idx + stride, per, slab);
        gates_finish<VEC>(a, slab, fb);
        if (two) gates_finish<VEC>(b, slab, fb);
    }
}

template <int VEC>
__global__ void __launch_bounds__(256)
gates_bwd_kernel(const float *__restrict__ conv, const float *__restrict__ state, const float *__restrict__ gns,
                 float *__restrict__ gconv, float *__restrict__ gstate, int B, long slab, float fb)
{
    const long per = slab / VEC;
    const long n = (long)B * per;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
        const long b = idx / per;
        const long e = (idx - b * per) * VEC;
        const float *cb = conv + b * 4 * slab + e;
        const float *sb = state + b * 2 * slab + e;
        const float *gb = gns + b * 2 * slab + e;
        float *gcb = gconv + b * 4 * slab + e;
        float *gsb = gstate + b * 2 * slab + e;
        float gi[VEC], gj[VEC], gf[VEC], go[VEC], c[VEC], gcn[VEC], ghn[VEC];
        float di[VEC], dj[VEC], df[VEC], dgo[VEC], dc[VEC], dz[VEC];
        if (VEC == 4) {
            *reinterpret_cast<float4 *>(gi) = ld_stream4(reinterpret_cast<const float4 *>(cb));
            *reinterpret_cast<float4 *>(gj) = ld_stream4(reinterpret_cast<const float4 *>(cb + slab));
            *reinterpret_cast<float4 *>(gf) = ld_stream4(reinterpret_cast<const float4 *>(cb + 2 * slab));
            *reinterpret_cast<float4 *>(go) = ld_stream4(reinterpret_cast<const float4 *>(cb + 3 * slab));
            *reinterpret_cast<float4 *>(c) = ld_stream4(reinterpret_cast<const float4 *>(sb));
            *reinterpret_cast<float4 *>(gcn) = ld_stream4(reinterpret_cast<const float4 *>(gb));
            *reinterpret_cast<float4 *>(ghn) = ld_stream4(reinterpret_cast<const float4 *>(gb + slab));
        } else {
            gi[0] = cb[0]; gj[0] = cb[slab]; gf[0] = cb[2 * slab]; go[0] = cb[3 * slab];
            c[0] = sb[0]; gcn[0] = gb[0]; ghn[0] = gb[slab];
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const float si = sigmoid_acc(gi[k]), sf = sigmoid_acc(gf[k] + fb), so = sigmoid_acc(go[k]);
            const float tj = tanh_acc(gj[k]);
            const float tc = tanh_acc(c[k] * sf + si * tj);
            const float gct = gcn[k] + ghn[k] * so * (1.f - tc * tc);
            di[k] = gct * tj * si * (1.f - si);
            dj[k] = gct * si * (1.f - tj * tj);
            df[k] = gct * c[k] * sf * (1.f - sf);
            dgo[k] = ghn[k] * tc * so * (1.f - so);
            dc[k] = gct * sf;
            dz[k] = 0.f;
        }
        if (VEC == 4) {
            *reinterpret_cast<float4 *>(gcb) = *reinterpret_cast<float4 *>(di);
            *reinterpret_cast<float4 *>(gcb + slab) = *reinterpret_cast<float4 *>(dj);
            *reinterpret_cast<float4 *>(gcb + 2 * slab) = *reinterpret_cast<float4 *>(df);
            *reinterpret_cast<float4 *>(gcb + 3 * slab) = *reinterpret_cast<float4 *>(dgo);
            *reinterpret_cast<float4 *>(gsb) = *reinterpret_cast<float4 *>(dc);
            *reinterpret_cast<float4 *>(gsb + slab) = *reinterpret_cast<float4 *>(dz);
        } else {
            gcb[0] = di[0]; gcb[slab] = dj[0]; gcb[2 * slab] = df[0]; gcb[3 * slab] = dgo[0];
            gsb[0] = dc[0]; gsb[slab] = 0.f;
        }
    }
}


template <int CT>
__global__ void __launch_bounds__(256)
warp_fwd_kernel(const float *__restrict__ img, const float *__restrict__ uv, float *__restrict__ out,
                int B, int C, const WarpGeom g)
{
    const int H = g.H, W = g.W;
    const int hw = H * W;
    const int n = B * hw;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int b = idx / hw, pix = idx - b * hw;
        const int y = pix / W, x = pix - y * W;
        const float *fl = uv + (long)b * 2 * hw + pix;
        const Taps t = make_taps(warp_coord(x, y, ld_stream(fl), ld_stream(fl + hw), g), H, W);
        const float *im = img + (long)b * C * hw;
        float *o = out + (long)b * C * hw + pix;
        TAI_CH_LOOP(CT, C) o[(long)ch * hw] = sample(im + (long)ch * hw, t, W);
    }
}

// Adjoint of the warp: image gradient scattered with red.global (the destination is zeroed by the launcher; the
// summation order of colliding taps is run-dependent), flow gradient written directly.  Tap offsets / validity /
// weights once per pixel, channel loop unrolled for C = 1 / 3 (CT == 0: runtime channel count).
template <int CT>
__global__ void __launch_bounds__(256)
warp_bwd_kernel(const float *__restrict__ img, const float *__restrict__ uv, const float *__restrict__ gout,
                float *__restrict__ gimg, float *__restrict__ guv, int B, int C, const WarpGeom geom)
{
    const int H = geom.H, W = geom.W;
    const int hw = H * W;
    const int n = B * hw;
    const float sx = (float)(W - 1) / (float)W, sy = (float)(H - 1) / (float)H;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int b = idx / hw, pix = idx - b * hw;
        const int y = pix / W, x = pix - y * W;
        const float *fl = uv + (long)b * 2 * hw + pix;
        const WarpCoord c = warp_coord(x, y, __ldg(fl), __ldg(fl + hw), geom);
        const Taps t = make_taps(c, H, W);
        const float x1 = (float)(c.x0 + 1), y1 = (float)(c.y0 + 1), x0 = (float)c.x0, y0 = (float)c.y0;
        const float ax = x1 - c.ix, bx = c.ix - x0, ay = y1 - c.iy, by = c.iy - y0;
        float gx = 0.f, gy = 0.f;
        const long base = (long)b * C * hw;
        TAI_CH_LOOP(CT, C) {
            const float *p = img + base + (long)ch * hw + t.o;
            const float g = __ldg(gout + base + (long)ch * hw + pix);
            const float nw = t.v00 ? __ldg(p) : 0.f, ne = t.v01 ? __ldg(p + 1) : 0.f;
            const float sw = t.v10 ? __ldg(p + W) : 0.f, se = t.v11 ? __ldg(p + W + 1) : 0.f;
            gx += g * ((ne - nw) * ay + (se - sw) * by);
            gy += g * ((sw - nw) * ax + (se - ne) * bx);
            if (gimg) {
                float *gi = gimg + base + (long)ch * hw + t.o;
                if (t.v00) atomicAdd(gi, g * ax * ay);
                if (t.v01) atomicAdd(gi + 1, g * bx * ay);
                if (t.v10) atomicAdd(gi + W, g * ax * by);
                if (t.v11) atomicAdd(gi + W + 1, g * bx * by);
            }
        }
        if (guv) {
            guv[(long)b * 2 * hw + pix] = gx * sx;
            guv[(long)b * 2 * hw + hw + pix] = gy * sy;
        }
    }
}

// slomo.py:312-316 in one pass: intermediate flows + both warps.
template <int CT>
__global__ void __launch_bounds__(256)
slomo_combine_warp_kernel(const float *__restrict__ i0, const float *__restrict__ i1,
                          const float *__restrict__ f01, const float *__restrict__ f10,
                          float c00, float c01, float c10, float c11,
                          float *__restrict__ ft0, float *__restrict__ ft1,
                          float *__restrict__ g0, float *__restrict__ g1, int B, int C, const WarpGeom g)
{
    const int H = g.H, W = g.W;
    const int hw = H * W;
    const int n = B * hw;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int b = idx / hw, pix = idx - b * hw;
        const int y = pix / W, x = pix - y * W;
        const long fu = (long)b * 2 * hw + pix, fv = fu + hw;
        const float a_u = ld_stream(f01 + fu), a_v = ld_stream(f01 + fv);
        const float b_u = ld_stream(f10 + fu), b_v = ld_stream(f10 + fv);
        // same association as the reference: (coef * F01) + (coef * F10), no contraction
        const float t0u = __fadd_rn(__fmul_rn(c00, a_u), __fmul_rn(c01, b_u));
        const float t0v = __fadd_rn(__fmul_rn(c00, a_v), __fmul_rn(c01, b_v));
        const float t1u = __fsub_rn(__fmul_rn(c10, a_u), __fmul_rn(c11, b_u));
        const float t1v = __fsub_rn(__fmul_rn(c10, a_v), __fmul_rn(c11, b_v));
        ft0[fu] = t0u; ft0[fv] = t0v;
        ft1[fu] = t1u; ft1[fv] = t1v;
        const Taps w0 = make_taps(warp_coord(x, y, t0u, t0v, g), H, W);
        const Taps w1 = make_taps(warp_coord(x, y, t1u, t1v, g), H, W);
        const long base = (long)b * C * hw;
        float r0[CT ? CT : 1], r1[CT ? CT : 1];
        if (CT) {  // all 8 * C gathers in flight before the first store
            TAI_CH_LOOP(CT, C) {
                r0[ch] = sample(i0 + base + (long)ch * hw, w0, W);
                r1[ch] = sample(i1 + base + (long)ch * hw, w1, W);
            }
            TAI_CH_LOOP(CT, C) {
                g0[base + (long)ch * hw + pix] = r0[ch];
                g1[base + (long)ch * hw + pix] = r1[ch];
            }
        } else {
            for (int ch = 0; ch < C; ++ch) {
                g0[base + (long)ch * hw + pix] = sample(i0 + base + (long)ch * hw, w0, W);
                g1[base + (long)ch * hw + pix] = sample(i1 + base + (long)ch * hw, w1, W);
            }
        }
    }
}

// slomo.py:320-328 in one pass: refine-add-clamp, two warps, visibility-weighted blend.
template <int CT>
__global__ void __launch_bounds__(256)
slomo_refine_blend_kernel(const float *__restrict__ i0, const float *__restrict__ i1,
                          const float *__restrict__ ft0, const float *__restrict__ ft1,
                          const float *__restrict__ d0, const float *__restrict__ d1,
                          const float *__restrict__ v0, float omt, float t,
                          float *__restrict__ out, int B, int C, const WarpGeom g)
{
    const int H = g.H, W = g.W;
    const int hw = H * W;
    const int n = B * hw;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int b = idx / hw, pix = idx - b * hw;
        const int y = pix / W, x = pix - y * W;
        const long fu = (long)b * 2 * hw + pix, fv = fu + hw;
        const float r0u = fminf(fmaxf(__fadd_rn(ld_stream(d0 + fu), ld_stream(ft0 + fu)), -1.f), 1.f);
        const float r0v = fminf(fmaxf(__fadd_rn(ld_stream(d0 + fv), ld_stream(ft0 + fv)), -1.f), 1.f);
        const float r1u = fminf(fmaxf(__fadd_rn(ld_stream(d1 + fu), ld_stream(ft1 + fu)), -1.f), 1.f);
        const float r1v = fminf(fmaxf(__fadd_rn(ld_stream(d1 + fv), ld_stream(ft1 + fv)), -1.f), 1.f);
        const Taps w0 = make_taps(warp_coord(x, y, r0u, r0v, g), H, W);
        const Taps w1 = make_taps(warp_coord(x, y, r1u, r1v, g), H, W);
        const float vis0 = ld_stream(v0 + (long)b * hw + pix);
        const float vis1 = 1.f - vis0;
        const float k0 = omt * vis0, k1 = t * vis1;
        const float inv = 1.f / (k0 + k1);   // one IEEE reciprocal for the C channels (the reference divides each: <= 1 ulp apart)
        const long base = (long)b * C * hw;
        if (CT) {
            float a0[CT ? CT : 1], a1[CT ? CT : 1];
            TAI_CH_LOOP(CT, C) {
                a0[ch] = sample(i0 + base + (long)ch * hw, w0, W);
                a1[ch] = sample(i1 + base + (long)ch * hw, w1, W);
            }
            TAI_CH_LOOP(CT, C) out[base + (long)ch * hw + pix] = (k0 * a0[ch] + k1 * a1[ch]) * inv;
        } else {
            for (int ch = 0; ch < C; ++ch) {
                const float a0 = sample(i0 + base + (long)ch * hw, w0, W), a1 = sample(i1 + base + (long)ch * hw, w1, W);
                out[base + (long)ch * hw + pix] = (k0 * a0 + k1 * a1) * inv;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Motion-stream prologue (tai.py:67-74; mcnet.py:439-447; util.py:22-41): frames in [-1, 1] -> [0, 1]
// (inverse_transform: (x + 1.) / 2) -> gray (bgr2gray: 0.1140 B + 0.5870 G + 0.2989 R, left to right; one channel:
// identity) -> temporal difference.  The reference spells this as 4-7 elementwise kernels per call over strided
// slices; here it is one pass.  One rounding per reference operation (no FMA contraction), so the result is
// bit-identical to the reference's FP32 chain.
template <int C>
__device__ __forceinline__ float gray01(const float *__restrict__ p, long cstride)
{
    if (C == 1) return __fmul_rn(__fadd_rn(ld_stream(p), 1.f), 0.5f);
    const float b = __fmul_rn(__fadd_rn(ld_stream(p), 1.f), 0.5f);
    const float g = __fmul_rn(__fadd_rn(ld_stream(p + cstride), 1.f), 0.5f);
    const float r = __fmul_rn(__fadd_rn(ld_stream(p + 2 * cstride), 1.f), 0.5f);
    return __fadd_rn(__fadd_rn(__fmul_rn(0.1140f, b), __fmul_rn(0.5870f, g)), __fmul_rn(0.2989f, r));
}

// frames [B,K,C,HW] -> out [B,K-1,1,HW]: out[b,k] = gray01(frame i(k+1)) - gray01(frame i(k)), i(k) = k or K-1-k
// (reverse: the time-reversed following frames of tai.py:71-74 without materialising the flip).  A thread walks
// the K frames of one pixel, so every frame is read once.
template <int C>
__global__ void __launch_bounds__(256)
gray_diff_frames_kernel(const float *__restrict__ frames, float *__restrict__ out, int B, int K, long hw, int reverse)
{
    const long n = (long)B * hw;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
        const long b = idx / hw, pix = idx - b * hw;
        const float *fb = frames + b * K * C * hw + pix;
        float prev = gray01<C>(fb + (long)(reverse ? K - 1 : 0) * C * hw, hw);
        for (int k = 1; k < K; ++k) {
            const float cur = gray01<C>(fb + (long)(reverse ? K - 1 - k : k) * C * hw, hw);
            out[(b * (K - 1) + (k - 1)) * hw + pix] = __fsub_rn(cur, prev);
            prev = cur;
        }
    }
}

// a, b [N,C,HW] -> out [N,1,HW] = gray01(a) - gray01(b)   (mcnet.py:439-447: the next motion input)
template <int C>
__global__ void __launch_bounds__(256)
gray_diff_pair_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ out, long N, long hw)
{
    const long n = N * hw;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
        const long s = idx / hw, pix = idx - s * hw;
        out[idx] = __fsub_rn(gray01<C>(a + s * C * hw + pix, hw), gray01<C>(b + s * C * hw + pix, hw));
    }
}

// adjoint: ga[c] = 0.5 w_c g, gb[c] = -0.5 w_c g (w = 1 for one channel); either output may be null
template <int C>
__global__ void __launch_bounds__(256)
gray_diff_pair_bwd_kernel(const float *__restrict__ gout, float *__restrict__ ga, float *__restrict__ gb, long N, long hw)
{
    const long n = N * hw;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
        const long s = idx / hw, pix = idx - s * hw;
        const float g = ld_stream(gout + idx);
        const float w[3] = {C == 1 ? 1.f : 0.1140f, 0.5870f, 0.2989f};
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float v = (g * w[c]) * 0.5f;
            if (ga) ga[(s * C + c) * hw + pix] = v;
            if (gb) gb[(s * C + c) * hw + pix] = -v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Output side of predict.py: frames in [-1, 1] -> 8-bit images (predict.py:124-134: clamp(-1, 1),
// (255 * ((x + 1.) / 2)).astype(uint8), BGR -> RGB for colour).  in [N,C,H,W] float -> out [N,H,W,C] bytes, the
// layout PIL / PNG encoders take; the device-to-host copy shrinks 4x.  FP32 arithmetic in the reference's order
// (the +1, the /2 and the *255 each round once; truncation toward zero): byte-exact.
__global__ void __launch_bounds__(256)
frames_to_u8_kernel(const float *__restrict__ in, unsigned char *__restrict__ out, long N, int C, long hw, int flip)
{
    const long n = N * hw;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
        const long b = idx / hw;
        const long pix = idx - b * hw;
        for (int c = 0; c < C; ++c) {
            const float x = ld_stream(in + (b * C + (flip ? C - 1 - c : c)) * hw + pix);
            const float v = fminf(fmaxf(x, -1.f), 1.f);
            const float s = __fmul_rn(255.f, __fmul_rn(__fadd_rn(v, 1.f), 0.5f));
            out[idx * C + c] = (unsigned char)(int)s;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// FFMA probe: 8 independent chains per thread, `iters` FMAs each (scalar FFMA or packed FFMA2).
template <bool PACKED>
__global__ void ffma_probe_kernel(float *sink, int iters)
{
    const float s = 1.0f + 1e-7f * (float)threadIdx.x, t = 1e-9f * (float)blockIdx.x;
    if (PACKED) {
        float2 a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = make_float2(0.1f * (float)k, 0.2f * (float)k);
        const float2 s2 = make_float2(s, s), t2 = make_float2(t, t);
#pragma unroll 1
        for (int it = 0; it < iters; it += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int k = 0; k < 4; ++k) a[k] = ffma2(a[k], s2, t2);
        }
        float r = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) r += a[k].x + a[k].y;
        sink[(long)blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else {
        float a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = 0.1f * (float)k;
#pragma unroll 1
        for (int it = 0; it < iters; it += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int k = 0; k < 8; ++k) a[k] = fmaf(a[k], s, t);
        }
        float r = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) r += a[k];
        sink[(long)blockIdx.x * blockDim.x + threadIdx.x] = r;
    }
}

}  // namespace tai

using namespace tai;

extern "C" int convlstm_gates_forward_b200(const float *conv_out, const float *state, float *new_state,
                                           int B, int F, int HW, float forget_bias, void *stream)
{
    TAI_REQUIRE(conv_out && state && new_state && B > 0 && F > 0 && HW > 0, TAI_ERR_INVALID_ARGUMENT,
                "convlstm_gates_forward_b200: bad arguments");
    const long slab = (long)F * HW;
    TAI_REQUIRE(fits_int31(4 * slab * B), TAI_ERR_TOO_LARGE, "convlstm_gates_forward_b200: tensor has >= 2^31 elements");
    const bool vec = (slab % 4 == 0) && ((((uintptr_t)conv_out | (uintptr_t)state | (uintptr_t)new_state) & 15) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    TimingScope ts("gates_fwd", st, 0.0, 28.0 * B * slab);  // read i,j,f,o,c; write c',h'
    if (vec) {
        gates_fwd_kernel<4><<<stream_grid_occ(gates_fwd_kernel<4>, B * slab / 4, 128, 2), 128, 0, st>>>(conv_out, state, new_state, B,
                                                                                                        slab, forget_bias);
    } else {
        gates_fwd_kernel<1><<<stream_grid_occ(gates_fwd_kernel<1>, B * slab, 128, 2), 128, 0, st>>>(conv_out, state, new_state, B, slab,
                                                                                                    forget_bias);
    }
    return check_launch("gates_fwd_kernel");
}

extern "C" int convlstm_gates_backward_b200(const float *conv_out, const float *state, const float *g_new_state,
                                            float *g_conv_out, float *g_state,
                                            int B, int F, int HW, float forget_bias, void *stream)
{
    TAI_REQUIRE(conv_out && state && g_new_state && g_conv_out && g_state && B > 0 && F > 0 && HW > 0,
                TAI_ERR_INVALID_ARGUMENT, "convlstm_gates_backward_b200: bad arguments");
    const long slab = (long)F * HW;
    TAI_REQUIRE(fits_int31(4 * slab * B), TAI_ERR_TOO_LARGE, "convlstm_gates_backward_b200: tensor has >= 2^31 elements");
    const bool vec = (slab % 4 == 0) &&
                     ((((uintptr_t)conv_out | (uintptr_t)state | (uintptr_t)g_new_state | (uintptr_t)g_conv_out |
                        (uintptr_t)g_state) & 15) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    TimingScope ts("gates_bwd", st, 0.0, 52.0 * B * slab);  // read i,j,f,o,c,gc',gh'; write 4 + 2
    if (vec)
        gates_bwd_kernel<4><<<stream_grid(B * slab / 4, 256), 256, 0, st>>>(conv_out, state, g_new_state, g_conv_out, g_state, B, slab, forget_bias);
    else
        gates_bwd_kernel<1><<<stream_grid(B * slab, 256), 256, 0, st>>>(conv_out, state, g_new_state, g_conv_out, g_state, B, slab, forget_bias);
    return check_launch("gates_bwd_kernel");
}

static int warp_args_ok(const char *who, int B, int C, int H, int W)
{
    TAI_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, TAI_ERR_INVALID_ARGUMENT, "%s: bad sizes B=%d C=%d H=%d W=%d", who, B, C, H, W);
    TAI_REQUIRE(fits_int31((long long)B * (C > 2 ? C : 2) * H * W), TAI_ERR_TOO_LARGE, "%s: tensor has >= 2^31 elements", who);
    return TAI_OK;
}

extern "C" int flow_warp_forward_b200(const float *img, const float *uv, float *out,
                                      int B, int C, int H, int W, void *stream)
{
    TAI_REQUIRE(img && uv && out, TAI_ERR_INVALID_ARGUMENT, "flow_warp_forward_b200: null pointer");
    int rc = warp_args_ok("flow_warp_forward_b200", B, C, H, W);
    if (rc) return rc;
    TimingScope ts("warp_fwd", (cudaStream_t)stream, 0.0, 4.0 * (2.0 + 2.0 * C) * B * H * W);
    const WarpGeom g = warp_geom(H, W);
    const unsigned grid = stream_grid((long)B * H * W, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 3)
        warp_fwd_kernel<3><<<grid, 256, 0, st>>>(img, uv, out, B, C, g);
    else if (C == 1)
        warp_fwd_kernel<1><<<grid, 256, 0, st>>>(img, uv, out, B, C, g);
    else
        warp_fwd_kernel<0><<<grid, 256, 0, st>>>(img, uv, out, B, C, g);
    return check_launch("warp_fwd_kernel");
}

extern "C" int flow_warp_backward_b200(const float *img, const float *uv, const float *grad_out,
                                       float *g_img, float *g_uv, int B, int C, int H, int W, void *stream)
{
    TAI_REQUIRE(img && uv && grad_out && (g_img || g_uv), TAI_ERR_INVALID_ARGUMENT, "flow_warp_backward_b200: null pointer");
    int rc = warp_args_ok("flow_warp_backward_b200", B, C, H, W);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (g_img) {
        cudaError_t e = cudaMemsetAsync(g_img, 0, sizeof(float) * (size_t)B * C * H * W, st);
        TAI_REQUIRE(e == cudaSuccess, TAI_ERR_CUDA, "flow_warp_backward_b200: memset: %s", cudaGetErrorString(e));
    }
    TimingScope ts("warp_bwd", st, 0.0, 4.0 * (4.0 + 3.0 * C) * B * H * W);
    const unsigned grid = stream_grid((long)B * H * W, 256);
    const WarpGeom g = warp_geom(H, W);
    if (C == 3)
        warp_bwd_kernel<3><<<grid, 256, 0, st>>>(img, uv, grad_out, g_img, g_uv, B, C, g);
    else if (C == 1)
        warp_bwd_kernel<1><<<grid, 256, 0, st>>>(img, uv, grad_out, g_img, g_uv, B, C, g);
    else
        warp_bwd_kernel<0><<<grid, 256, 0, st>>>(img, uv, grad_out, g_img, g_uv, B, C, g);
    return check_launch("warp_bwd_kernel");
}

extern "C" int slomo_flow_combine_warp_forward_b200(const float *i0, const float *i1,
                                                    const float *f01, const float *f10, double t,
                                                    float *f_t0, float *f_t1, float *g0, float *g1,
                                                    int B, int C, int H, int W, void *stream)
{
    TAI_REQUIRE(i0 && i1 && f01 && f10 && f_t0 && f_t1 && g0 && g1, TAI_ERR_INVALID_ARGUMENT,
                "slomo_flow_combine_warp_forward_b200: null pointer");
    int rc = warp_args_ok("slomo_flow_combine_warp_forward_b200", B, C, H, W);
    if (rc) return rc;
    // Python-float (double) scalars of slomo.py:313-314, rounded to FP32 when they meet the tensor.
    const float c00 = (float)(-(1.0 - t) * t), c01 = (float)(t * t);
    const float c10 = (float)((1.0 - t) * (1.0 - t)), c11 = (float)(t * (1.0 - t));
    TimingScope ts("slomo_combine_warp", (cudaStream_t)stream, 0.0, 4.0 * (8.0 + 4.0 * C) * B * H * W);
    const WarpGeom g = warp_geom(H, W);
    const unsigned grid = stream_grid((long)B * H * W, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 3)
        slomo_combine_warp_kernel<3><<<grid, 256, 0, st>>>(i0, i1, f01, f10, c00, c01, c10, c11, f_t0, f_t1, g0, g1, B, C, g);
    else if (C == 1)
        slomo_combine_warp_kernel<1><<<grid, 256, 0, st>>>(i0, i1, f01, f10, c00, c01, c10, c11, f_t0, f_t1, g0, g1, B, C, g);
    else
        slomo_combine_warp_kernel<0><<<grid, 256, 0, st>>>(i0, i1, f01, f10, c00, c01, c10, c11, f_t0, f_t1, g0, g1, B, C, g);
    return check_launch("slomo_combine_warp_kernel");
}

extern "C" int slomo_refine_blend_forward_b200(const float *i0, const float *i1,
                                               const float *f_t0, const float *f_t1,
                                               const float *d_t0, const float *d_t1, const float *v_t0,
                                               double t, float *out, int B, int C, int H, int W, void *stream)
{
    TAI_REQUIRE(i0 && i1 && f_t0 && f_t1 && d_t0 && d_t1 && v_t0 && out, TAI_ERR_INVALID_ARGUMENT,
                "slomo_refine_blend_forward_b200: null pointer");
    int rc = warp_args_ok("slomo_refine_blend_forward_b200", B, C, H, W);
    if (rc) return rc;
    TimingScope ts("slomo_refine_blend", (cudaStream_t)stream, 0.0, 4.0 * (9.0 + 3.0 * C) * B * H * W);
    const WarpGeom g = warp_geom(H, W);
    const unsigned grid = stream_grid((long)B * H * W, 256);
    cudaStream_t st = (cudaStream_t)stream;
    const float omt = (float)(1.0 - t), ft = (float)t;
    if (C == 3)
        slomo_refine_blend_kernel<3><<<grid, 256, 0, st>>>(i0, i1, f_t0, f_t1, d_t0, d_t1, v_t0, omt, ft, out, B, C, g);
    else if (C == 1)
        slomo_refine_blend_kernel<1><<<grid, 256, 0, st>>>(i0, i1, f_t0, f_t1, d_t0, d_t1, v_t0, omt, ft, out, B, C, g);
    else
        slomo_refine_blend_kernel<0><<<grid, 256, 0, st>>>(i0, i1, f_t0, f_t1, d_t0, d_t1, v_t0, omt, ft, out, B, C, g);
    return check_launch("slomo_refine_blend_kernel");
}

extern "C" int gray_difference_frames_b200(const float *frames, float *out, int B, int K, int C, int H, int W, int reverse,
                                           void *stream)
{
    TAI_REQUIRE(frames && out && B > 0 && K > 1 && (C == 1 || C == 3) && H > 0 && W > 0, TAI_ERR_INVALID_ARGUMENT,
                "gray_difference_frames_b200: bad arguments (B=%d K=%d C=%d H=%d W=%d; C must be 1 or 3, K >= 2)", B, K, C, H, W);
    TAI_REQUIRE(fits_int31((long long)B * K * C * H * W), TAI_ERR_TOO_LARGE, "gray_difference_frames_b200: tensor has >= 2^31 elements");
    cudaStream_t st = (cudaStream_t)stream;
    const long hw = (long)H * W;
    TimingScope ts("gray_diff_frames", st, 0.0, 4.0 * ((double)K * C + (K - 1)) * B * hw);
    const unsigned grid = stream_grid((long)B * hw, 256);
    if (C == 3)
        gray_diff_frames_kernel<3><<<grid, 256, 0, st>>>(frames, out, B, K, hw, reverse ? 1 : 0);
    else
        gray_diff_frames_kernel<1><<<grid, 256, 0, st>>>(frames, out, B, K, hw, reverse ? 1 : 0);
    return check_launch("gray_diff_frames_kernel");
}

extern "C" int gray_difference_pair_forward_b200(const float *a, const float *b, float *out, long long N, int C, int H, int W,
                                                 void *stream)
{
    TAI_REQUIRE(a && b && out && N > 0 && (C == 1 || C == 3) && H > 0 && W > 0, TAI_ERR_INVALID_ARGUMENT,
                "gray_difference_pair_forward_b200: bad arguments (C must be 1 or 3)");
    TAI_REQUIRE(fits_int31(N * C * H * W), TAI_ERR_TOO_LARGE, "gray_difference_pair_forward_b200: tensor has >= 2^31 elements");
    cudaStream_t st = (cudaStream_t)stream;
    const long hw = (long)H * W;
    TimingScope ts("gray_diff_pair", st, 0.0, 4.0 * (2.0 * C + 1.0) * N * hw);
    const unsigned grid = stream_grid((long)N * hw, 256);
    if (C == 3)
        gray_diff_pair_kernel<3><<<grid, 256, 0, st>>>(a, b, out, (long)N, hw);
    else
        gray_diff_pair_kernel<1><<<grid, 256, 0, st>>>(a, b, out, (long)N, hw);
    return check_launch("gray_diff_pair_kernel");
}

extern "C" int gray_difference_pair_backward_b200(const float *grad_out, float *g_a, float *g_b, long long N, int C, int H, int W,
                                                  void *stream)
{
    TAI_REQUIRE(grad_out && (g_a || g_b) && N > 0 && (C == 1 || C == 3) && H > 0 && W > 0, TAI_ERR_INVALID_ARGUMENT,
                "gray_difference_pair_backward_b200: bad arguments (C must be 1 or 3)");
    TAI_REQUIRE(fits_int31(N * C * H * W), TAI_ERR_TOO_LARGE, "gray_difference_pair_backward_b200: tensor has >= 2^31 elements");
    cudaStream_t st = (cudaStream_t)stream;
    const long hw = (long)H * W;
    TimingScope ts("gray_diff_pair_bwd", st, 0.0, 4.0 * (1.0 + ((g_a ? 1.0 : 0.0) + (g_b ? 1.0 : 0.0)) * C) * N * hw);
    const unsigned grid = stream_grid((long)N * hw, 256);
    if (C == 3)
        gray_diff_pair_bwd_kernel<3><<<grid, 256, 0, st>>>(grad_out, g_a, g_b, (long)N, hw);
    else
        gray_diff_pair_bwd_kernel<1><<<grid, 256, 0, st>>>(grad_out, g_a, g_b, (long)N, hw);
    return check_launch("gray_diff_pair_bwd_kernel");
}

extern "C" int frames_to_uint8_b200(const float *frames, unsigned char *out, long long N, int C, int H, int W, int flip_channels,
                                    void *stream)
{
    TAI_REQUIRE(frames && out && N > 0 && C > 0 && H > 0 && W > 0, TAI_ERR_INVALID_ARGUMENT, "frames_to_uint8_b200: bad arguments");
    TAI_REQUIRE(fits_int31(N * (long long)C * H * W), TAI_ERR_TOO_LARGE, "frames_to_uint8_b200: tensor has >= 2^31 elements");
    cudaStream_t st = (cudaStream_t)stream;
    TimingScope ts("frames_to_u8", st, 0.0, 5.0 * N * C * H * W);
    frames_to_u8_kernel<<<stream_grid((long)N * H * W, 256), 256, 0, st>>>(frames, out, (long)N, C, (long)H * W, flip_channels ? 1 : 0);
    return check_launch("frames_to_u8_kernel");
}

extern "C" int tai_b200_ffma_probe(float *sink, int grid, int block, int iters, int packed, void *stream)
{
    TAI_REQUIRE(sink && grid > 0 && block > 0 && block <= 1024 && iters > 0, TAI_ERR_INVALID_ARGUMENT,
                "tai_b200_ffma_probe: bad arguments");
    if (packed)
        ffma_probe_kernel<true><<<grid, block, 0, (cudaStream_t)stream>>>(sink, iters);
    else
        ffma_probe_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(sink, iters);
    return check_launch("ffma_probe_kernel");
}
