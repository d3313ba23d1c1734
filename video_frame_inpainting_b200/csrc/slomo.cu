// Super SloMo intermediate-frame stages (slomo.py:307-340) for sm_100a, ALL T time steps per launch, with their
// adjoints.
//
// The reference runs, per middle frame t: two scalar-tensor flow combinations (slomo.py:313-314), two
// FlowWarper calls (315-316: host meshgrid + H2D copy + ~8 elementwise kernels + grid_sample each), a torch.cat of
// six tensors (318), the refinement U-Net, two clamps, two more FlowWarper calls and ~10 elementwise kernels for the
// visibility blend (320-328).  The T time steps do not depend on each other (the U-Nets carry no state), so here
// they are one batch:
//
//   slomo_interp_input   F_t0, F_t1 for every t, both warps, and the refinement network's input tensor
//                        X[t*B+b] = cat(I0, g(I0,F_t0), F_t0, F_t1, g(I1,F_t1), I1) written in place (no cat
//                        copy); the flows also go to the model's F_t_*_collector outputs in the reference's
//                        (reversed) time order, which is where the blend stage reads them;
//   slomo_refine_blend   refine-add-clamp, two warps, visibility-weighted blend, result written straight into
//                        pred[B,T,C,H,W] (reversed time order, slomo.py:332-340);
//   *_bwd                the adjoints w.r.t. the flows / refinements / visibility: pure gathers (the warp's flow
//                        gradient needs the same four taps as the forward), one thread per pixel accumulating
//                        over the T time steps -- no atomics, deterministic.  Image gradients (I0 / I1 are
//                        network inputs) are not produced here; a caller that needs them takes the composed route
//                        through flow_warp_backward_b200.
//
// All four are HBM-bound gather / stream kernels; at T*B = 24 UCF frames a launch moves 150-250 MB instead of
// the 20-50 MB of the per-t kernels, which is what lifts them from launch-latency-bound to bandwidth-bound.
#include "common.cuh"
#include "warp.cuh"

namespace tai {

constexpr int kSlomoMaxT = 16;

// Python-float (double) scalars of slomo.py:312-314,325-328, rounded to FP32 where they meet a tensor.
struct SlomoTimes {
    int T;
    float c00[kSlomoMaxT], c01[kSlomoMaxT], c10[kSlomoMaxT], c11[kSlomoMaxT], omt[kSlomoMaxT], t[kSlomoMaxT];
};

static SlomoTimes slomo_times(int T)
{
    SlomoTimes tm;
    tm.T = T;
    for (int i = 0; i < kSlomoMaxT; ++i) {
        const double t = (i + 1.0) / (T + 1.0);  // t = (t_ + 1) / (T + 1), true division (slomo.py:2,312)
        tm.c00[i] = (float)(-(1.0 - t) * t);
        tm.c01[i] = (float)(t * t);
        tm.c10[i] = (float)((1.0 - t) * (1.0 - t));
        tm.c11[i] = (float)(t * (1.0 - t));
        tm.omt[i] = (float)(1.0 - t);
        tm.t[i] = (float)t;
    }
    return tm;
}

struct TFlows {
    float t0u, t0v, t1u, t1v;
};

// same association as the reference: (coef * F01) + (coef * F10), one rounding per operation, no contraction
__device__ __forceinline__ TFlows combine_flows(const SlomoTimes &tm, int t, float a_u, float a_v, float b_u, float b_v)
{
    TFlows f;
    f.t0u = __fadd_rn(__fmul_rn(tm.c00[t], a_u), __fmul_rn(tm.c01[t], b_u));
    f.t0v = __fadd_rn(__fmul_rn(tm.c00[t], a_v), __fmul_rn(tm.c01[t], b_v));
    f.t1u = __fsub_rn(__fmul_rn(tm.c10[t], a_u), __fmul_rn(tm.c11[t], b_u));
    f.t1v = __fsub_rn(__fmul_rn(tm.c10[t], a_v), __fmul_rn(tm.c11[t], b_v));
    return f;
}

// The four taps of a sample point and the derivative of the bilinear interpolant w.r.t. the sampling position.
struct TapVals {
    float val, ddx, ddy;
};

__device__ __forceinline__ TapVals sample_with_derivative(const float *__restrict__ plane, const Taps &t, int W,
                                                          float ax, float bx, float ay, float by)
{
    const float *p = plane + t.o;
    const float nw = t.v00 ? __ldg(p) : 0.f, ne = t.v01 ? __ldg(p + 1) : 0.f;
    const float sw = t.v10 ? __ldg(p + W) : 0.f, se = t.v11 ? __ldg(p + W + 1) : 0.f;
    TapVals r;
    r.val = nw * t.wnw + ne * t.wne + sw * t.wsw + se * t.wse;
    r.ddx = (ne - nw) * ay + (se - sw) * by;
    r.ddy = (sw - nw) * ax + (se - ne) * bx;
    return r;
}

struct Frac {
    float ax, bx, ay, by;  // x1 - ix, ix - x0, y1 - iy, iy - y0
};
__device__ __forceinline__ Frac frac_of(const WarpCoord &c)
{
    Frac f;
    f.ax = (float)(c.x0 + 1) - c.ix;
    f.bx = c.ix - (float)c.x0;
    f.ay = (float)(c.y0 + 1) - c.iy;
    f.by = c.iy - (float)c.y0;
    return f;
}

// ------------------------------------------------------------------------------------------------
// slomo.py:312-318 for every t: thread = (b, pixel), loop over the T time steps (the flows and the two direct
// image pixels are loaded once).
template <int CT>
__global__ void __launch_bounds__(256)
slomo_interp_input_kernel(const float *__restrict__ i0, const float *__restrict__ i1,
                          const float *__restrict__ f01, const float *__restrict__ f10, const SlomoTimes tm,
                          float *__restrict__ X, float *__restrict__ ft0c, float *__restrict__ ft1c,
                          int B, int C, const WarpGeom g)
{
    const int H = g.H, W = g.W, T = tm.T;
    const int hw = H * W;
    const int n = B * hw;
    const int XC = 4 * C + 4;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int b = idx / hw, pix = idx - b * hw;
        const int y = pix / W, x = pix - y * W;
        const long fu = (long)b * 2 * hw + pix, fv = fu + hw;
        const float a_u = ld_stream(f01 + fu), a_v = ld_stream(f01 + fv);
        const float b_u = ld_stream(f10 + fu), b_v = ld_stream(f10 + fv);
        const long base = (long)b * C * hw;
        float p0[CT ? CT : 1], p1[CT ? CT : 1];
        if (CT) {
            TAI_CH_LOOP(CT, C) {
                p0[ch] = __ldg(i0 + base + (long)ch * hw + pix);
                p1[ch] = __ldg(i1 + base + (long)ch * hw + pix);
            }
        }
        for (int t = 0; t < T; ++t) {
            const TFlows f = combine_flows(tm, t, a_u, a_v, b_u, b_v);
            const Taps w0 = make_taps(warp_coord(x, y, f.t0u, f.t0v, g), H, W);
            const Taps w1 = make_taps(warp_coord(x, y, f.t1u, f.t1v, g), H, W);
            float *xo = X + ((long)t * B + b) * XC * hw + pix;
            const long fo = ((long)b * T + (T - 1 - t)) * 2 * hw + pix;   // collectors: reversed time order
            if (CT) {
                float r0[CT ? CT : 1], r1[CT ? CT : 1];
                TAI_CH_LOOP(CT, C) {   // all gathers in flight before the first store
                    r0[ch] = sample(i0 + base + (long)ch * hw, w0, W);
                    r1[ch] = sample(i1 + base + (long)ch * hw, w1, W);
                }
                TAI_CH_LOOP(CT, C) {
                    xo[(long)ch * hw] = p0[ch];
                    xo[(long)(C + ch) * hw] = r0[ch];
                    xo[(long)(2 * C + 4 + ch) * hw] = r1[ch];
                    xo[(long)(3 * C + 4 + ch) * hw] = p1[ch];
                }
            } else {
                for (int ch = 0; ch < C; ++ch) {
                    xo[(long)ch * hw] = __ldg(i0 + base + (long)ch * hw + pix);
                    xo[(long)(C + ch) * hw] = sample(i0 + base + (long)ch * hw, w0, W);
                    xo[(long)(2 * C + 4 + ch) * hw] = sample(i1 + base + (long)ch * hw, w1, W);
                    xo[(long)(3 * C + 4 + ch) * hw] = __ldg(i1 + base + (long)ch * hw + pix);
                }
            }
            xo[(long)(2 * C) * hw] = f.t0u;
            xo[(long)(2 * C + 1) * hw] = f.t0v;
            xo[(long)(2 * C + 2) * hw] = f.t1u;
            xo[(long)(2 * C + 3) * hw] = f.t1v;
            ft0c[fo] = f.t0u;
            ft0c[fo + hw] = f.t0v;
            ft1c[fo] = f.t1u;
            ft1c[fo + hw] = f.t1v;
        }
    }
}

// Adjoint w.r.t. F_0_1 and F_1_0: for every t the gradient reaching F_t0 is (flow gradient of the warp of I0)
// + (gradient of X's F_t0 channels) + (gradient of the collector), likewise F_t1; the linear combination of
// slomo.py:313-314 is transposed and summed over t in registers.
template <int CT>
__global__ void __launch_bounds__(256)
slomo_interp_input_bwd_kernel(const float *__restrict__ i0, const float *__restrict__ i1,
                              const float *__restrict__ f01, const float *__restrict__ f10, const SlomoTimes tm,
                              const float *__restrict__ gX, const float *__restrict__ gft0c,
                              const float *__restrict__ gft1c, float *__restrict__ gf01, float *__restrict__ gf10,
                              int B, int C, const WarpGeom g)
{
    const int H = g.H, W = g.W, T = tm.T;
    const int hw = H * W;
    const int n = B * hw;
    const int XC = 4 * C + 4;
    const float sx = (float)(W - 1) / (float)W, sy = (float)(H - 1) / (float)H;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int b = idx / hw, pix = idx - b * hw;
        const int y = pix / W, x = pix - y * W;
        const long fu = (long)b * 2 * hw + pix, fv = fu + hw;
        const float a_u = ld_stream(f01 + fu), a_v = ld_stream(f01 + fv);
        const float b_u = ld_stream(f10 + fu), b_v = ld_stream(f10 + fv);
        const long base = (long)b * C * hw;
        float gau = 0.f, gav = 0.f, gbu = 0.f, gbv = 0.f;
        // streamed operands of one time step: the gradients of X's warped-frame and flow channels and of the collectors
        struct StepIn {
            float go0[CT ? CT : 1], go1[CT ? CT : 1], xf[4], cf[4];
        };
        auto load_step = [&](int t, StepIn &in) {
            const float *gxo = gX + ((long)t * B + b) * XC * hw + pix;
            if (CT) {
                TAI_CH_LOOP(CT, C) {
                    in.go0[ch] = ld_stream(gxo + (long)(C + ch) * hw);
                    in.go1[ch] = ld_stream(gxo + (long)(2 * C + 4 + ch) * hw);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) in.xf[k] = ld_stream(gxo + (long)(2 * C + k) * hw);
            const long fo = ((long)b * T + (T - 1 - t)) * 2 * hw + pix;
            in.cf[0] = gft0c ? ld_stream(gft0c + fo) : 0.f;
            in.cf[1] = gft0c ? ld_stream(gft0c + fo + hw) : 0.f;
            in.cf[2] = gft1c ? ld_stream(gft1c + fo) : 0.f;
            in.cf[3] = gft1c ? ld_stream(gft1c + fo + hw) : 0.f;
        };
        StepIn cur;
        load_step(0, cur);
        for (int t = 0; t < T; ++t) {
            StepIn nxt = cur;
            if (t + 1 < T) load_step(t + 1, nxt);   // in flight under this step's gathers (the kernel was latency-bound:
                                                    // three dependent round trips per pixel, issue slots 33 % busy)
            const TFlows f = combine_flows(tm, t, a_u, a_v, b_u, b_v);
            const WarpCoord c0 = warp_coord(x, y, f.t0u, f.t0v, g), c1 = warp_coord(x, y, f.t1u, f.t1v, g);
            const Taps w0 = make_taps(c0, H, W), w1 = make_taps(c1, H, W);
            const Frac q0 = frac_of(c0), q1 = frac_of(c1);
            const float *gxo = gX + ((long)t * B + b) * XC * hw + pix;
            float gx0 = 0.f, gy0 = 0.f, gx1 = 0.f, gy1 = 0.f;
            TAI_CH_LOOP(CT, C) {
                const float go0 = CT ? cur.go0[CT ? ch : 0] : ld_stream(gxo + (long)(C + ch) * hw);
                const float go1 = CT ? cur.go1[CT ? ch : 0] : ld_stream(gxo + (long)(2 * C + 4 + ch) * hw);
                const TapVals s0 = sample_with_derivative(i0 + base + (long)ch * hw, w0, W, q0.ax, q0.bx, q0.ay, q0.by);
                const TapVals s1 = sample_with_derivative(i1 + base + (long)ch * hw, w1, W, q1.ax, q1.bx, q1.ay, q1.by);
                gx0 += go0 * s0.ddx;
                gy0 += go0 * s0.ddy;
                gx1 += go1 * s1.ddx;
                gy1 += go1 * s1.ddy;
            }
            const float g0u = gx0 * sx + cur.xf[0] + cur.cf[0];
            const float g0v = gy0 * sy + cur.xf[1] + cur.cf[1];
            const float g1u = gx1 * sx + cur.xf[2] + cur.cf[2];
            const float g1v = gy1 * sy + cur.xf[3] + cur.cf[3];
            // F_t0 = c00 F01 + c01 F10 ; F_t1 = c10 F01 - c11 F10
            gau += tm.c00[t] * g0u + tm.c10[t] * g1u;
            gav += tm.c00[t] * g0v + tm.c10[t] * g1v;
            gbu += tm.c01[t] * g0u - tm.c11[t] * g1u;
            gbv += tm.c01[t] * g0v - tm.c11[t] * g1v;
            cur = nxt;
        }
        gf01[fu] = gau;
        gf01[fv] = gav;
        gf10[fu] = gbu;
        gf10[fv] = gbv;
    }
}

// ------------------------------------------------------------------------------------------------
// slomo.py:320-328 for every (t, b): thread = (b, pixel), loop over t with the NEXT time step's nine streamed
// operands already in flight (a thread per (t, b, pixel) ran at 45 % of the HBM peak: two dependent memory round
// trips -- streamed operands, then the gathers they address -- per thread and only ~24 resident warps per SM to
// hide them; the loop overlaps the first trip of t+1 with the gathers and stores of t).
__device__ __forceinline__ float clamp_pm1(float v) { return fminf(fmaxf(v, -1.f), 1.f); }

struct RefineIn {
    float d0u, d0v, d1u, d1v, f0u, f0v, f1u, f1v, vis;
};

__device__ __forceinline__ RefineIn refine_load(const float *__restrict__ ft0c, const float *__restrict__ ft1c,
                                                const float *__restrict__ d0, const float *__restrict__ d1,
                                                const float *__restrict__ v0, int t, int b, int pix, int B, int T, int hw)
{
    const long nb = (long)t * B + b;
    const long du = nb * 2 * hw + pix, fu = ((long)b * T + (T - 1 - t)) * 2 * hw + pix;  // collectors: reversed time
    RefineIn r;
    r.d0u = ld_stream(d0 + du); r.d0v = ld_stream(d0 + du + hw);
    r.d1u = ld_stream(d1 + du); r.d1v = ld_stream(d1 + du + hw);
    r.f0u = ld_stream(ft0c + fu); r.f0v = ld_stream(ft0c + fu + hw);
    r.f1u = ld_stream(ft1c + fu); r.f1v = ld_stream(ft1c + fu + hw);
    r.vis = ld_stream(v0 + nb * hw + pix);
    return r;
}

template <int CT>
__global__ void __launch_bounds__(256, 4)
slomo_refine_blend_t_kernel(const float *__restrict__ i0, const float *__restrict__ i1,
                            const float *__restrict__ ft0c, const float *__restrict__ ft1c,
                            const float *__restrict__ d0, const float *__restrict__ d1, const float *__restrict__ v0,
                            const SlomoTimes tm, float *__restrict__ pred, int B, int C, const WarpGeom g)
{
    const int H = g.H, W = g.W, T = tm.T;
    const int hw = H * W;
    const int n = B * hw;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int b = idx / hw, pix = idx - b * hw;
        const int y = pix / W, x = pix - y * W;
        const long base = (long)b * C * hw;
        RefineIn cur = refine_load(ft0c, ft1c, d0, d1, v0, 0, b, pix, B, T, hw);
        for (int t = 0; t < T; ++t) {
            RefineIn nxt = cur;
            if (t + 1 < T) nxt = refine_load(ft0c, ft1c, d0, d1, v0, t + 1, b, pix, B, T, hw);
            const float r0u = clamp_pm1(__fadd_rn(cur.d0u, cur.f0u)), r0v = clamp_pm1(__fadd_rn(cur.d0v, cur.f0v));
            const float r1u = clamp_pm1(__fadd_rn(cur.d1u, cur.f1u)), r1v = clamp_pm1(__fadd_rn(cur.d1v, cur.f1v));
            const Taps w0 = make_taps(warp_coord(x, y, r0u, r0v, g), H, W);
            const Taps w1 = make_taps(warp_coord(x, y, r1u, r1v, g), H, W);
            const float vis1 = 1.f - cur.vis;
            const float k0 = tm.omt[t] * cur.vis, k1 = tm.t[t] * vis1;
            const float inv = 1.f / (k0 + k1);   // one IEEE reciprocal for the C channels (the reference divides each)
            float *out = pred + ((long)b * T + (T - 1 - t)) * C * hw + pix;   // reversed time order (slomo.py:332-340)
            if (CT) {
                float a0[CT ? CT : 1], a1[CT ? CT : 1];
                TAI_CH_LOOP(CT, C) {
                    a0[ch] = sample(i0 + base + (long)ch * hw, w0, W);
                    a1[ch] = sample(i1 + base + (long)ch * hw, w1, W);
                }
                TAI_CH_LOOP(CT, C) out[(long)ch * hw] = (k0 * a0[ch] + k1 * a1[ch]) * inv;
            } else {
                for (int ch = 0; ch < C; ++ch) {
                    const float a0 = sample(i0 + base + (long)ch * hw, w0, W), a1 = sample(i1 + base + (long)ch * hw, w1, W);
                    out[(long)ch * hw] = (k0 * a0 + k1 * a1) * inv;
                }
            }
            cur = nxt;
        }
    }
}

// Adjoint w.r.t. the refined flows (the clamp passes gradient on [-1, 1] inclusive, as torch.clamp does; dF_t and
// F_t enter through their sum, so they receive the same gradient) and the visibility map:
//   out = (k0 a0 + k1 a1) / n,  k0 = (1-t) V,  k1 = t (1 - V),  n = k0 + k1
//   d out / d V = ((1-t) a0 - t a1 - out ((1-t) - t)) / n
template <int CT>
__global__ void __launch_bounds__(256)
slomo_refine_blend_t_bwd_kernel(const float *__restrict__ i0, const float *__restrict__ i1,
                                const float *__restrict__ ft0c, const float *__restrict__ ft1c,
                                const float *__restrict__ d0, const float *__restrict__ d1,
                                const float *__restrict__ v0, const SlomoTimes tm, const float *__restrict__ gpred,
                                float *__restrict__ gft0c, float *__restrict__ gft1c, float *__restrict__ gd0,
                                float *__restrict__ gd1, float *__restrict__ gv0, int B, int C, const WarpGeom g)
{
    const int H = g.H, W = g.W, T = tm.T;
    const int hw = H * W;
    const int n = B * hw;
    const float sx = (float)(W - 1) / (float)W, sy = (float)(H - 1) / (float)H;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int b = idx / hw, pix = idx - b * hw;
        const int y = pix / W, x = pix - y * W;
        const long base = (long)b * C * hw;
        RefineIn cur = refine_load(ft0c, ft1c, d0, d1, v0, 0, b, pix, B, T, hw);
        for (int t = 0; t < T; ++t) {
            RefineIn nxt = cur;
            if (t + 1 < T) nxt = refine_load(ft0c, ft1c, d0, d1, v0, t + 1, b, pix, B, T, hw);
            const long nb = (long)t * B + b;
            const long du = nb * 2 * hw + pix, dv = du + hw;
            const long slot = (long)b * T + (T - 1 - t);
            const long fu = slot * 2 * hw + pix, fv = fu + hw;
            const float s0u = __fadd_rn(cur.d0u, cur.f0u), s0v = __fadd_rn(cur.d0v, cur.f0v);
            const float s1u = __fadd_rn(cur.d1u, cur.f1u), s1v = __fadd_rn(cur.d1v, cur.f1v);
            const WarpCoord c0 = warp_coord(x, y, clamp_pm1(s0u), clamp_pm1(s0v), g);
            const WarpCoord c1 = warp_coord(x, y, clamp_pm1(s1u), clamp_pm1(s1v), g);
            const Taps w0 = make_taps(c0, H, W), w1 = make_taps(c1, H, W);
            const Frac q0 = frac_of(c0), q1 = frac_of(c1);
            const float omt = tm.omt[t], tt = tm.t[t];
            const float k0 = omt * cur.vis, k1 = tt * (1.f - cur.vis);
            const float inv = 1.f / (k0 + k1);
            const float *go = gpred + slot * C * hw + pix;
            float gx0 = 0.f, gy0 = 0.f, gx1 = 0.f, gy1 = 0.f, gv = 0.f;
            TAI_CH_LOOP(CT, C) {
                const float gch = ld_stream(go + (long)ch * hw);
                const TapVals a0 = sample_with_derivative(i0 + base + (long)ch * hw, w0, W, q0.ax, q0.bx, q0.ay, q0.by);
                const TapVals a1 = sample_with_derivative(i1 + base + (long)ch * hw, w1, W, q1.ax, q1.bx, q1.ay, q1.by);
                const float ga0 = gch * k0 * inv, ga1 = gch * k1 * inv;
                gx0 += ga0 * a0.ddx;
                gy0 += ga0 * a0.ddy;
                gx1 += ga1 * a1.ddx;
                gy1 += ga1 * a1.ddy;
                const float o = (k0 * a0.val + k1 * a1.val) * inv;
                gv += gch * ((omt * a0.val - tt * a1.val) - o * (omt - tt)) * inv;
            }
            const float r0u = (s0u >= -1.f && s0u <= 1.f) ? gx0 * sx : 0.f, r0v = (s0v >= -1.f && s0v <= 1.f) ? gy0 * sy : 0.f;
            const float r1u = (s1u >= -1.f && s1u <= 1.f) ? gx1 * sx : 0.f, r1v = (s1v >= -1.f && s1v <= 1.f) ? gy1 * sy : 0.f;
            gd0[du] = r0u; gd0[dv] = r0v;
            gd1[du] = r1u; gd1[dv] = r1v;
            gft0c[fu] = r0u; gft0c[fv] = r0v;
            gft1c[fu] = r1u; gft1c[fv] = r1v;
            gv0[nb * hw + pix] = gv;
            cur = nxt;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The same two kernels for W % 4 == 0 and C in {1, 3}: FOUR adjacent pixels per thread and the frames staged in
// shared memory.  The per-pixel kernels above are issue-bound, not latency-bound (ncu: 25 M warp instructions for
// 2 M pixel-steps = 400 per pixel-step, issue slots 57 % busy, DRAM 25 %): 64-bit address arithmetic for nine
// streamed loads, 24 predicated gathers and three stores per pixel-step outweighs the ~120 FP instructions of
// the coordinate chains and the blend.  With four pixels per thread the streamed operands move as 128-bit
// accesses (a quarter of the address arithmetic).  The refined flows are clamped to [-1, 1] pixel (slomo.py:
// 320-321) and ix = (x + u)(W-1)/W, so the four taps of pixel (x, y) lie in columns x-2..x+1 and rows y-2..y+1
// whatever the inputs: a CTA that owns a 32 x 16 pixel tile stages a 35 x 19 patch of I0 and of I1 per channel
// ONCE (zero-filled outside the image: the sampler's zero padding, so taps need no validity predicates) and
// uses it for all T steps.  Warp = 4 rows x 8 threads; patch pitch 37 makes the bank of a tap 4 tx + 5 r + const:
// conflict-free when the lanes' displacements agree.  floor() and the tap weights come from the same FP32
// coordinate chain as everywhere else; only the address of a tap is formed relative to the patch.
constexpr int kQW = 32, kQH = 16, kQLo = 2, kQPW = kQW + 3, kQPH = kQH + 3, kQPitch = 37, kQThreads = 128;

template <int CT>
__device__ __forceinline__ void refine_stage_patch(float *patch, const float *__restrict__ i0, const float *__restrict__ i1,
                                                   int b, int tx0, int ty0, int H, int W)
{
    // Same thread -> (row, quad) mapping as the compute phase: a 128-bit load per (plane, row, quad) for the 32
    // interior columns, the quads at the tile's edges add the two left and the one right halo columns.  Loads go
    // to clamped (always valid) addresses and are zeroed by selects: no divergent branches around them.
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tx = lane & 7, r = lane >> 3;
    const int hw = H * W;
    const int x = tx0 + 4 * tx;
    const bool okx = x < W;                                   // W % 4 == 0
    const bool left = tx == 0 && tx0 > 0, right = tx == 7 && tx0 + kQW < W;
    const float *b0 = i0 + (long)b * CT * hw, *b1 = i1 + (long)b * CT * hw;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int ry = pass * 16 + 4 * warp + r;
        if (ry < kQPH) {                                      // second pass: patch rows 16..18 only
            const int gy = ty0 - kQLo + ry;
            const bool ok = (unsigned)gy < (unsigned)H && okx;
            const int off = min(max(gy, 0), H - 1) * W + (okx ? x : 0);
            float4 v[2 * CT];
            float2 vl[2 * CT];
            float vr[2 * CT];
#pragma unroll
            for (int plane = 0; plane < 2 * CT; ++plane) {
                const float *src = (plane >= CT ? b1 + (plane - CT) * hw : b0 + plane * hw) + off;
                v[plane] = __ldg(reinterpret_cast<const float4 *>(src));
                vl[plane] = make_float2(0.f, 0.f);
                vr[plane] = 0.f;
                if (left) vl[plane] = __ldg(reinterpret_cast<const float2 *>(src - 2));
                if (right) vr[plane] = __ldg(src + 4);
            }
            float *dst = patch + ry * kQPitch + kQLo + 4 * tx;
#pragma unroll
            for (int plane = 0; plane < 2 * CT; ++plane) {
                float *d = dst + plane * (kQPH * kQPitch);
                d[0] = ok ? v[plane].x : 0.f; d[1] = ok ? v[plane].y : 0.f;
                d[2] = ok ? v[plane].z : 0.f; d[3] = ok ? v[plane].w : 0.f;
                if (tx == 0) { d[-2] = ok ? vl[plane].x : 0.f; d[-1] = ok ? vl[plane].y : 0.f; }
                if (tx == 7) d[4] = ok ? vr[plane] : 0.f;
            }
        }
    }
}

struct RefineIn4 {
    float4 d0u, d0v, d1u, d1v, f0u, f0v, f1u, f1v, vis;
};

__device__ __forceinline__ RefineIn4 refine_load4(const float *__restrict__ ft0c, const float *__restrict__ ft1c,
                                                  const float *__restrict__ d0, const float *__restrict__ d1,
                                                  const float *__restrict__ v0, int t, int b, int pix, int B, int T, int hw)
{
    const long nb = (long)t * B + b;
    const long du = nb * 2 * hw + pix, fu = ((long)b * T + (T - 1 - t)) * 2 * hw + pix;  // collectors: reversed time
    auto L = [](const float *p) { return ld_stream4(reinterpret_cast<const float4 *>(p)); };
    RefineIn4 r;
    r.d0u = L(d0 + du); r.d0v = L(d0 + du + hw);
    r.d1u = L(d1 + du); r.d1v = L(d1 + du + hw);
    r.f0u = L(ft0c + fu); r.f0v = L(ft0c + fu + hw);
    r.f1u = L(ft1c + fu); r.f1v = L(ft1c + fu + hw);
    r.vis = L(v0 + nb * hw + pix);
    return r;
}

__device__ __forceinline__ float comp(const float4 &v, int j) { return j == 0 ? v.x : j == 1 ? v.y : j == 2 ? v.z : v.w; }
__device__ __forceinline__ void set_comp(float4 &v, int j, float a)
{
    if (j == 0) v.x = a; else if (j == 1) v.y = a; else if (j == 2) v.z = a; else v.w = a;
}

// offset of the north-west tap inside a patch plane (the clamp only guards the address against inputs that cannot
// occur: |flow| <= 1 keeps every tap inside the patch)
__device__ __forceinline__ int patch_offset(const WarpCoord &c, int tx0, int ty0)
{
    const int lx = min(max(c.x0 - (tx0 - kQLo), 0), kQPW - 2), ly = min(max(c.y0 - (ty0 - kQLo), 0), kQPH - 2);
    return ly * kQPitch + lx;
}

struct QuadGeom {
    int tx0, ty0, b, x, y, pix;
    bool active;
};

__device__ __forceinline__ QuadGeom quad_geom(int tiles_x, int tiles_y, int H, int W)
{
    QuadGeom q;
    int tile = blockIdx.x;
    q.tx0 = (tile % tiles_x) * kQW;
    tile /= tiles_x;
    q.ty0 = (tile % tiles_y) * kQH;
    q.b = tile / tiles_y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    q.x = q.tx0 + 4 * (lane & 7);
    q.y = q.ty0 + 4 * warp + (lane >> 3);
    q.active = q.x < W && q.y < H;      // W % 4 == 0: a quad is inside or outside as a whole
    q.pix = q.active ? q.y * W + q.x : 0;
    return q;
}

template <int CT>
__global__ void __launch_bounds__(kQThreads, 4)
slomo_refine_blend_quad_kernel(const float *__restrict__ i0, const float *__restrict__ i1,
                               const float *__restrict__ ft0c, const float *__restrict__ ft1c,
                               const float *__restrict__ d0, const float *__restrict__ d1, const float *__restrict__ v0,
                               const SlomoTimes tm, float *__restrict__ pred, int B, int tiles_x, int tiles_y, const WarpGeom g)
{
    __shared__ float patch[2 * CT * kQPH * kQPitch];
    constexpr int PLANE = kQPH * kQPitch;
    const int H = g.H, W = g.W, T = tm.T;
    const int hw = H * W;
    const QuadGeom q = quad_geom(tiles_x, tiles_y, H, W);
    RefineIn4 ra = refine_load4(ft0c, ft1c, d0, d1, v0, 0, q.b, q.pix, B, T, hw);   // in flight while the patch is staged
    refine_stage_patch<CT>(patch, i0, i1, q.b, q.tx0, q.ty0, H, W);
    __syncthreads();
    if (!q.active) return;
    auto step = [&](const RefineIn4 &cur, int t) {
        const float omt = tm.omt[t], tt = tm.t[t];
        float4 o[CT];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float r0u = clamp_pm1(__fadd_rn(comp(cur.d0u, j), comp(cur.f0u, j)));
            const float r0v = clamp_pm1(__fadd_rn(comp(cur.d0v, j), comp(cur.f0v, j)));
            const float r1u = clamp_pm1(__fadd_rn(comp(cur.d1u, j), comp(cur.f1u, j)));
            const float r1v = clamp_pm1(__fadd_rn(comp(cur.d1v, j), comp(cur.f1v, j)));
            const WarpCoord c0 = warp_coord(q.x + j, q.y, r0u, r0v, g), c1 = warp_coord(q.x + j, q.y, r1u, r1v, g);
            const Frac q0 = frac_of(c0), q1 = frac_of(c1);
            const float *p0 = patch + patch_offset(c0, q.tx0, q.ty0), *p1 = patch + CT * PLANE + patch_offset(c1, q.tx0, q.ty0);
            const float vis = comp(cur.vis, j);
            const float k0 = omt * vis, k1 = tt * (1.f - vis);
            const float inv = 1.f / (k0 + k1);
            // same weights and summation order as sample(): nw*wnw + ne*wne + sw*wsw + se*wse
            const float w00 = q0.ax * q0.ay, w01 = q0.bx * q0.ay, w02 = q0.ax * q0.by, w03 = q0.bx * q0.by;
            const float w10 = q1.ax * q1.ay, w11 = q1.bx * q1.ay, w12 = q1.ax * q1.by, w13 = q1.bx * q1.by;
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) {
                const float *a = p0 + ch * PLANE, *c = p1 + ch * PLANE;
                const float a0 = a[0] * w00 + a[1] * w01 + a[kQPitch] * w02 + a[kQPitch + 1] * w03;
                const float a1 = c[0] * w10 + c[1] * w11 + c[kQPitch] * w12 + c[kQPitch + 1] * w13;
                set_comp(o[ch], j, (k0 * a0 + k1 * a1) * inv);
            }
        }
        float *out = pred + ((long)q.b * T + (T - 1 - t)) * CT * hw + q.pix;
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) *reinterpret_cast<float4 *>(out + (long)ch * hw) = o[ch];
    };
    // two steps per trip, the operands of the next step in flight while this one is computed (no register copies)
    for (int t = 0; t < T; t += 2) {
        RefineIn4 rb;
        if (t + 1 < T) rb = refine_load4(ft0c, ft1c, d0, d1, v0, t + 1, q.b, q.pix, B, T, hw);
        step(ra, t);
        if (t + 1 < T) {
            if (t + 2 < T) ra = refine_load4(ft0c, ft1c, d0, d1, v0, t + 2, q.b, q.pix, B, T, hw);
            step(rb, t + 1);
        }
    }
}

template <int CT>
__global__ void __launch_bounds__(kQThreads)
slomo_refine_blend_quad_bwd_kernel(const float *__restrict__ i0, const float *__restrict__ i1,
                                   const float *__restrict__ ft0c, const float *__restrict__ ft1c,
                                   const float *__restrict__ d0, const float *__restrict__ d1,
                                   const float *__restrict__ v0, const SlomoTimes tm, const float *__restrict__ gpred,
                                   float *__restrict__ gft0c, float *__restrict__ gft1c, float *__restrict__ gd0,
                                   float *__restrict__ gd1, float *__restrict__ gv0, int B, int tiles_x, int tiles_y,
                                   const WarpGeom g)
{
    __shared__ float patch[2 * CT * kQPH * kQPitch];
    constexpr int PLANE = kQPH * kQPitch;
    const int H = g.H, W = g.W, T = tm.T;
    const int hw = H * W;
    const float sx = (float)(W - 1) / (float)W, sy = (float)(H - 1) / (float)H;
    const QuadGeom q = quad_geom(tiles_x, tiles_y, H, W);
    RefineIn4 cur = refine_load4(ft0c, ft1c, d0, d1, v0, 0, q.b, q.pix, B, T, hw);
    refine_stage_patch<CT>(patch, i0, i1, q.b, q.tx0, q.ty0, H, W);
    __syncthreads();
    if (!q.active) return;
    for (int t = 0; t < T; ++t) {
        const long nb = (long)t * B + q.b;
        const long du = nb * 2 * hw + q.pix, dv = du + hw;
        const long slot = (long)q.b * T + (T - 1 - t);
        const long fu = slot * 2 * hw + q.pix, fv = fu + hw;
        const float *go = gpred + slot * CT * hw + q.pix;
        float4 gch[CT];
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) gch[ch] = ld_stream4(reinterpret_cast<const float4 *>(go + (long)ch * hw));
        RefineIn4 nxt = cur;
        if (t + 1 < T) nxt = refine_load4(ft0c, ft1c, d0, d1, v0, t + 1, q.b, q.pix, B, T, hw);
        const float omt = tm.omt[t], tt = tm.t[t];
        float4 o0u, o0v, o1u, o1v, ov;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float s0u = __fadd_rn(comp(cur.d0u, j), comp(cur.f0u, j)), s0v = __fadd_rn(comp(cur.d0v, j), comp(cur.f0v, j));
            const float s1u = __fadd_rn(comp(cur.d1u, j), comp(cur.f1u, j)), s1v = __fadd_rn(comp(cur.d1v, j), comp(cur.f1v, j));
            const WarpCoord c0 = warp_coord(q.x + j, q.y, clamp_pm1(s0u), clamp_pm1(s0v), g);
            const WarpCoord c1 = warp_coord(q.x + j, q.y, clamp_pm1(s1u), clamp_pm1(s1v), g);
            const Frac q0 = frac_of(c0), q1 = frac_of(c1);
            const float *p0 = patch + patch_offset(c0, q.tx0, q.ty0), *p1 = patch + CT * PLANE + patch_offset(c1, q.tx0, q.ty0);
            const float vis = comp(cur.vis, j);
            const float k0 = omt * vis, k1 = tt * (1.f - vis);
            const float inv = 1.f / (k0 + k1);
            float gx0 = 0.f, gy0 = 0.f, gx1 = 0.f, gy1 = 0.f, gv = 0.f;
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) {
                const float gc = comp(gch[ch], j);
                const float *a = p0 + ch * PLANE, *c = p1 + ch * PLANE;
                const float nw0 = a[0], ne0 = a[1], sw0 = a[kQPitch], se0 = a[kQPitch + 1];
                const float nw1 = c[0], ne1 = c[1], sw1 = c[kQPitch], se1 = c[kQPitch + 1];
                const float a0 = nw0 * (q0.ax * q0.ay) + ne0 * (q0.bx * q0.ay) + sw0 * (q0.ax * q0.by) + se0 * (q0.bx * q0.by);
                const float a1 = nw1 * (q1.ax * q1.ay) + ne1 * (q1.bx * q1.ay) + sw1 * (q1.ax * q1.by) + se1 * (q1.bx * q1.by);
                const float ga0 = gc * k0 * inv, ga1 = gc * k1 * inv;
                gx0 += ga0 * ((ne0 - nw0) * q0.ay + (se0 - sw0) * q0.by);
                gy0 += ga0 * ((sw0 - nw0) * q0.ax + (se0 - ne0) * q0.bx);
                gx1 += ga1 * ((ne1 - nw1) * q1.ay + (se1 - sw1) * q1.by);
                gy1 += ga1 * ((sw1 - nw1) * q1.ax + (se1 - ne1) * q1.bx);
                const float o = (k0 * a0 + k1 * a1) * inv;
                gv += gc * ((omt * a0 - tt * a1) - o * (omt - tt)) * inv;
            }
            set_comp(o0u, j, (s0u >= -1.f && s0u <= 1.f) ? gx0 * sx : 0.f);
            set_comp(o0v, j, (s0v >= -1.f && s0v <= 1.f) ? gy0 * sy : 0.f);
            set_comp(o1u, j, (s1u >= -1.f && s1u <= 1.f) ? gx1 * sx : 0.f);
            set_comp(o1v, j, (s1v >= -1.f && s1v <= 1.f) ? gy1 * sy : 0.f);
            set_comp(ov, j, gv);
        }
        auto S = [](float *p, const float4 &v) { *reinterpret_cast<float4 *>(p) = v; };
        S(gd0 + du, o0u); S(gd0 + dv, o0v);
        S(gd1 + du, o1u); S(gd1 + dv, o1v);
        S(gft0c + fu, o0u); S(gft0c + fv, o0v);
        S(gft1c + fu, o1u); S(gft1c + fv, o1v);
        S(gv0 + nb * hw + q.pix, ov);
        cur = nxt;
    }
}

static int slomo_args_ok(const char *who, int B, int T, int C, int H, int W)
{
    TAI_REQUIRE(B > 0 && T > 0 && C > 0 && H > 0 && W > 0, TAI_ERR_INVALID_ARGUMENT,
                "%s: bad sizes B=%d T=%d C=%d H=%d W=%d", who, B, T, C, H, W);
    TAI_REQUIRE(T <= kSlomoMaxT, TAI_ERR_UNSUPPORTED, "%s: T=%d middle frames per launch (limit %d)", who, T, kSlomoMaxT);
    TAI_REQUIRE(fits_int31((long long)B * T * (4 * C + 4) * H * W), TAI_ERR_TOO_LARGE, "%s: tensor has >= 2^31 elements", who);
    return TAI_OK;
}

}  // namespace tai

using namespace tai;

// grid = one balanced wave of the kernel's real residency (equal grid-stride trips per thread)
#define TAI_SLOMO_DISPATCH(KERNEL, ITEMS, ST, ...)                                                    \
    do {                                                                                              \
        if (C == 3)                                                                                   \
            KERNEL<3><<<stream_grid_occ(KERNEL<3>, ITEMS, 256, 1), 256, 0, ST>>>(__VA_ARGS__);        \
        else if (C == 1)                                                                              \
            KERNEL<1><<<stream_grid_occ(KERNEL<1>, ITEMS, 256, 1), 256, 0, ST>>>(__VA_ARGS__);        \
        else                                                                                          \
            KERNEL<0><<<stream_grid_occ(KERNEL<0>, ITEMS, 256, 1), 256, 0, ST>>>(__VA_ARGS__);        \
    } while (0)

extern "C" int slomo_interp_input_forward_b200(const float *i0, const float *i1, const float *f01, const float *f10,
                                               float *interp_input, float *f_t0_collector, float *f_t1_collector,
                                               int B, int T, int C, int H, int W, void *stream)
{
    TAI_REQUIRE(i0 && i1 && f01 && f10 && interp_input && f_t0_collector && f_t1_collector, TAI_ERR_INVALID_ARGUMENT,
                "slomo_interp_input_forward_b200: null pointer");
    int rc = slomo_args_ok("slomo_interp_input_forward_b200", B, T, C, H, W);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // per (b, pixel): read 4 flows + 2C direct pixels; per t: 2C gathered pixels, write 4C + 4 (X) + 4 (collectors)
    TimingScope ts("slomo_interp_input", st, 0.0, 4.0 * (4.0 + 2.0 * C + T * (6.0 * C + 8.0)) * B * H * W);
    const WarpGeom g = warp_geom(H, W);
    const long items = (long)B * H * W;
    const SlomoTimes tm = slomo_times(T);
    TAI_SLOMO_DISPATCH(slomo_interp_input_kernel, items, st, i0, i1, f01, f10, tm, interp_input, f_t0_collector,
                       f_t1_collector, B, C, g);
    return check_launch("slomo_interp_input_kernel");
}

extern "C" int slomo_interp_input_backward_b200(const float *i0, const float *i1, const float *f01, const float *f10,
                                                const float *g_interp_input, const float *g_f_t0_collector,
                                                const float *g_f_t1_collector, float *g_f01, float *g_f10,
                                                int B, int T, int C, int H, int W, void *stream)
{
    TAI_REQUIRE(i0 && i1 && f01 && f10 && g_interp_input && g_f01 && g_f10, TAI_ERR_INVALID_ARGUMENT,
                "slomo_interp_input_backward_b200: null pointer");
    int rc = slomo_args_ok("slomo_interp_input_backward_b200", B, T, C, H, W);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const double coll = (g_f_t0_collector ? 2.0 : 0.0) + (g_f_t1_collector ? 2.0 : 0.0);
    TimingScope ts("slomo_interp_input_bwd", st, 0.0, 4.0 * (8.0 + T * (4.0 * C + 4.0 + coll)) * B * H * W);
    const WarpGeom g = warp_geom(H, W);
    const long items = (long)B * H * W;
    const SlomoTimes tm = slomo_times(T);
    TAI_SLOMO_DISPATCH(slomo_interp_input_bwd_kernel, items, st, i0, i1, f01, f10, tm, g_interp_input, g_f_t0_collector,
                       g_f_t1_collector, g_f01, g_f10, B, C, g);
    return check_launch("slomo_interp_input_bwd_kernel");
}

extern "C" int slomo_refine_blend_batched_forward_b200(const float *i0, const float *i1, const float *f_t0_collector,
                                                       const float *f_t1_collector, const float *d_t0,
                                                       const float *d_t1, const float *v_t0, float *pred,
                                                       int B, int T, int C, int H, int W, void *stream)
{
    TAI_REQUIRE(i0 && i1 && f_t0_collector && f_t1_collector && d_t0 && d_t1 && v_t0 && pred, TAI_ERR_INVALID_ARGUMENT,
                "slomo_refine_blend_batched_forward_b200: null pointer");
    int rc = slomo_args_ok("slomo_refine_blend_batched_forward_b200", B, T, C, H, W);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // compulsory bytes: nine streamed planes and C output planes per step, the two frames once per clip
    TimingScope ts("slomo_refine_blend_t", st, 0.0, 4.0 * ((9.0 + C) * T + 2.0 * C) * B * H * W);
    const WarpGeom g = warp_geom(H, W);
    const long items = (long)B * H * W;
    const SlomoTimes tm = slomo_times(T);
    const int tiles_x = ceil_div(W, kQW), tiles_y = ceil_div(H, kQH);
    const bool quads_aligned = (((uintptr_t)f_t0_collector | (uintptr_t)f_t1_collector | (uintptr_t)d_t0 | (uintptr_t)d_t1 |
                                 (uintptr_t)v_t0 | (uintptr_t)pred) & 15) == 0;
    if ((C == 1 || C == 3) && W % 4 == 0 && quads_aligned && fits_int31((long long)B * tiles_x * tiles_y)) {   // four pixels per thread
        const unsigned grid = (unsigned)((long)B * tiles_x * tiles_y);
        if (C == 3)
            slomo_refine_blend_quad_kernel<3><<<grid, kQThreads, 0, st>>>(i0, i1, f_t0_collector, f_t1_collector, d_t0, d_t1, v_t0, tm,
                                                                     pred, B, tiles_x, tiles_y, g);
        else
            slomo_refine_blend_quad_kernel<1><<<grid, kQThreads, 0, st>>>(i0, i1, f_t0_collector, f_t1_collector, d_t0, d_t1, v_t0, tm,
                                                                     pred, B, tiles_x, tiles_y, g);
        return check_launch("slomo_refine_blend_quad_kernel");
    }
    TAI_SLOMO_DISPATCH(slomo_refine_blend_t_kernel, items, st, i0, i1, f_t0_collector, f_t1_collector, d_t0, d_t1, v_t0, tm,
                       pred, B, C, g);
    return check_launch("slomo_refine_blend_t_kernel");
}

extern "C" int slomo_refine_blend_batched_backward_b200(const float *i0, const float *i1, const float *f_t0_collector,
                                                        const float *f_t1_collector, const float *d_t0,
                                                        const float *d_t1, const float *v_t0, const float *g_pred,
                                                        float *g_f_t0_collector, float *g_f_t1_collector, float *g_d_t0,
                                                        float *g_d_t1, float *g_v_t0,
                                                        int B, int T, int C, int H, int W, void *stream)
{
    TAI_REQUIRE(i0 && i1 && f_t0_collector && f_t1_collector && d_t0 && d_t1 && v_t0 && g_pred && g_f_t0_collector &&
                    g_f_t1_collector && g_d_t0 && g_d_t1 && g_v_t0,
                TAI_ERR_INVALID_ARGUMENT, "slomo_refine_blend_batched_backward_b200: null pointer");
    int rc = slomo_args_ok("slomo_refine_blend_batched_backward_b200", B, T, C, H, W);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    TimingScope ts("slomo_refine_blend_t_bwd", st, 0.0, 4.0 * ((18.0 + C) * T + 2.0 * C) * B * H * W);
    const WarpGeom g = warp_geom(H, W);
    const long items = (long)B * H * W;
    const SlomoTimes tm = slomo_times(T);
    const int tiles_x = ceil_div(W, kQW), tiles_y = ceil_div(H, kQH);
    const bool quads_aligned = (((uintptr_t)f_t0_collector | (uintptr_t)f_t1_collector | (uintptr_t)d_t0 | (uintptr_t)d_t1 |
                                 (uintptr_t)v_t0 | (uintptr_t)g_pred | (uintptr_t)g_f_t0_collector | (uintptr_t)g_f_t1_collector |
                                 (uintptr_t)g_d_t0 | (uintptr_t)g_d_t1 | (uintptr_t)g_v_t0) & 15) == 0;
    if ((C == 1 || C == 3) && W % 4 == 0 && quads_aligned && fits_int31((long long)B * tiles_x * tiles_y)) {
        const unsigned grid = (unsigned)((long)B * tiles_x * tiles_y);
        if (C == 3)
            slomo_refine_blend_quad_bwd_kernel<3><<<grid, kQThreads, 0, st>>>(i0, i1, f_t0_collector, f_t1_collector, d_t0, d_t1, v_t0, tm,
                                                                         g_pred, g_f_t0_collector, g_f_t1_collector, g_d_t0,
                                                                         g_d_t1, g_v_t0, B, tiles_x, tiles_y, g);
        else
            slomo_refine_blend_quad_bwd_kernel<1><<<grid, kQThreads, 0, st>>>(i0, i1, f_t0_collector, f_t1_collector, d_t0, d_t1, v_t0, tm,
                                                                         g_pred, g_f_t0_collector, g_f_t1_collector, g_d_t0,
                                                                         g_d_t1, g_v_t0, B, tiles_x, tiles_y, g);
        return check_launch("slomo_refine_blend_quad_bwd_kernel");
    }
    TAI_SLOMO_DISPATCH(slomo_refine_blend_t_bwd_kernel, items, st, i0, i1, f_t0_collector, f_t1_collector, d_t0, d_t1, v_t0,
                       tm, g_pred, g_f_t0_collector, g_f_t1_collector, g_d_t0, g_d_t1, g_v_t0, B, C, g);
    return check_launch("slomo_refine_blend_t_bwd_kernel");
}
