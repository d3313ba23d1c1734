/*
 * tai_b200.h -- C ABI of the B200-native TAI / bi-TAI hot path (libtai_b200.so).
 *
 * Every entry point is `extern "C"`, takes plain device pointers, sizes and a CUDA stream
 * (passed as void* == cudaStream_t; NULL = legacy default stream), borrows all pointers, fully
 * overwrites its outputs (callers may pass uninitialised memory -- the reference's zero-fill,
 * SeparableConvolution.py:36,69-71, is not needed), keeps no state between calls, performs no
 * allocation and no host synchronisation (CUDA-graph capturable) and returns 0 on success or a
 * negative TAI_ERR_* code; it never throws.  tai_b200_last_error() returns a thread-local
 * message for the last failure.
 *
 * All tensors are FP32, contiguous NCHW, and live on the current CUDA device.  Element counts
 * are validated to stay below 2^31 per tensor (same limit as the reference, kernel.cu:172).
 *
 * Each declaration names the reference interface it replaces; paths are relative to the
 * reference repository (MichiganCOG/video-frame-inpainting).
 */
#ifndef TAI_B200_H
#define TAI_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define TAI_B200_ABI_VERSION 1

enum {
    TAI_OK = 0,
    TAI_ERR_INVALID_ARGUMENT = -1, /* null pointer, non-positive size, even ks for a fused-pad entry, ... */
    TAI_ERR_UNSUPPORTED = -2,      /* ks or C outside the compiled range (ks <= 64)                        */
    TAI_ERR_TOO_LARGE = -3,        /* a tensor would have >= 2^31 elements                                 */
    TAI_ERR_CUDA = -4              /* a CUDA runtime call / launch failed; see tai_b200_last_error()       */
};

int tai_b200_abi_version(void);
const char *tai_b200_last_error(void);

/* Number of kernels this library has launched so far in this process (monotonic counter,
 * used by bench.py for its "gpu_launches" claim). */
long long tai_b200_launch_count(void);

/* Which kernel family the last separable-convolution launch of this thread took: "fwd:v3" / "fwd:v5" (persistent,
 * TMA-fed), "fwd:tiled" (LDG-fed, any ks <= 64: shapes TMA cannot describe, W % 4 != 0 or a misaligned base),
 * "fwd:simple" (shape-agnostic); "bwd_vh:*", "bwd_i:*" likewise.  Static string; for the op sweep's report. */
const char *tai_b200_last_path(void);

/* Measurement support for bench.py: when enabled every kernel launch of this library is bracketed by
 * CUDA events on its own stream; the report is a JSON array of
 * {"name","launches","ms","flops","bytes"} (ms = summed event time, flops/bytes = the ALGORITHMIC work
 * the launcher attributes to those launches).  Enabling clears earlier records. */
int tai_b200_timing_enable(int on);
int tai_b200_timing_report(char *buf, int buflen);

/* ------------------------------------------------------------------------------------------
 * Per-pixel separable local convolution.
 *
 * Replaces  int SeparableConvolution_cuda_forward(THCudaTensor* input, vertical, horizontal,
 *                                                 output, int ks)
 *           src/separable_convolution/cfile/SeparableConvolution_cuda.h:1-7  (-> kernel.cu:164-185,19-47)
 *
 *   input      [B, C, Hi, Wi]            vertical, horizontal [B, ks, Ho, Wo]
 *   output     [B, C, Ho, Wo]            Ho = Hi-ks+1, Wo = Wi-ks+1
 *   output[b,c,y,x] = sum_i sum_j input[b,c,y+i,x+j] * vertical[b,i,y,x] * horizontal[b,j,y,x]
 */
int SeparableConvolution_cuda_forward_b200(const float *input, const float *vertical,
                                           const float *horizontal, float *output,
                                           int B, int C, int Hi, int Wi, int ks, void *stream);

/* Replaces  int SeparableConvolution_cuda_backward(THCudaTensor* grad_output, input, vertical,
 *                 horizontal, grad_input, grad_vertical, grad_horizontal, int ks)
 *           src/separable_convolution/cfile/SeparableConvolution_cuda.h:9-18 (-> kernel.cu:187-242,49-162)
 *
 *   grad_output [B,C,Ho,Wo] -> grad_input [B,C,Hi,Wi], grad_vertical / grad_horizontal [B,ks,Ho,Wo].
 *   Any of the three gradient pointers may be NULL to skip that gradient.
 *   Determinism: grad_vertical / grad_horizontal are bit-reproducible.  grad_input is accumulated with FP32
 *   red.global from overlapping source tiles (the callee zeroes it on `stream` first), so its summation order -- and
 *   with it the last bits -- can differ from run to run, where the reference's gather kernel (kernel.cu:120-162) is
 *   deterministic; the difference stays inside the 1e-4 parity bound, and integer-valued inputs (the tap-count test)
 *   are exact in any order.  The same holds for g_img of flow_warp_backward_b200 and for the fused
 *   tai_fused_backward_b200 (g_pred_f / g_pred_b).
 */
int SeparableConvolution_cuda_backward_b200(const float *grad_output, const float *input,
                                            const float *vertical, const float *horizontal,
                                            float *grad_input, float *grad_vertical,
                                            float *grad_horizontal,
                                            int B, int C, int Hi, int Wi, int ks, void *stream);

/* ------------------------------------------------------------------------------------------
 * Fused TAI call site: replication pad + both separable convolutions + blend.
 *
 * Replaces the op sequence  modulePad -> separableConvolution (x2)   src/models/tai/tai.py:170-171,229-236
 *                           0.5*Dot1 + 0.5*Dot2                      src/models/tai/tai.py:105
 *                           (1-w)*Dot1 + w*Dot2                      src/models/twi/twi.py:105
 *
 *   pred_f, pred_b [B,C,H,W] (UNPADDED);  v1,h1,v2,h2 [B,ks,H,W];  ks odd, p = ks/2
 *   dot1 = sepconv(reppad(pred_f,p), v1, h1);  dot2 = sepconv(reppad(pred_b,p), v2, h2)
 *   pred = a*dot1 + b*dot2.   dot1 / dot2 may be NULL (not stored).
 */
int tai_fused_forward_b200(const float *pred_f, const float *pred_b,
                           const float *v1, const float *h1, const float *v2, const float *h2,
                           float *pred, float *dot1, float *dot2,
                           int B, int C, int H, int W, int ks, float a, float b, void *stream);

/* Backward of tai_fused_forward_b200.  grad_pred / grad_dot1 / grad_dot2 [B,C,H,W], any may be
 * NULL (treated as zero) but not all three.  The effective upstream gradients are
 * gD1 = a*grad_pred + grad_dot1, gD2 = b*grad_pred + grad_dot2.  Outputs: gradients w.r.t. the
 * UNPADDED predictions (replication-pad adjoint folded in) and the four kernel maps.
 * `workspace` must hold at least tai_fused_backward_workspace_bytes(...) bytes. */
long long tai_fused_backward_workspace_bytes(int B, int C, int H, int W, int ks);
int tai_fused_backward_b200(const float *grad_pred, const float *grad_dot1, const float *grad_dot2,
                            const float *pred_f, const float *pred_b,
                            const float *v1, const float *h1, const float *v2, const float *h2,
                            float *g_pred_f, float *g_pred_b,
                            float *g_v1, float *g_h1, float *g_v2, float *g_h2,
                            void *workspace,
                            int B, int C, int H, int W, int ks, float a, float b, void *stream);

/* Standalone replication pad and its adjoint (torch.nn.ReplicationPad2d(p), tai.py:170-171). */
int replication_pad_forward_b200(const float *in, float *out, int N, int H, int W, int p, void *stream);
int replication_pad_backward_b200(const float *grad_out, float *grad_in, int N, int H, int W, int p, void *stream);

/* ------------------------------------------------------------------------------------------
 * ConvLSTM gate path.
 *
 * Replaces the elementwise chain of ConvLstmCell.forward, src/models/mcnet/mcnet.py:287-293:
 *   c,h = chunk(state,2,1); (i,j,f,o) = chunk(conv_out,4,1)
 *   c' = c*sigmoid(f+forget_bias) + sigmoid(i)*tanh(j);  h' = tanh(c')*sigmoid(o)
 *   new_state = cat(c',h')
 *   conv_out [B,4F,h,w], state [B,2F,h,w] -> new_state [B,2F,h,w];  HW = h*w.
 */
int convlstm_gates_forward_b200(const float *conv_out, const float *state, float *new_state,
                                int B, int F, int HW, float forget_bias, void *stream);
/* g_new_state [B,2F,HW] -> g_conv_out [B,4F,HW], g_state [B,2F,HW] (h half written as zero). */
int convlstm_gates_backward_b200(const float *conv_out, const float *state, const float *g_new_state,
                                 float *g_conv_out, float *g_state,
                                 int B, int F, int HW, float forget_bias, void *stream);

/* ------------------------------------------------------------------------------------------
 * Super SloMo bilinear backward warp.
 *
 * Replaces FlowWarper.forward, src/models/slomo/slomo.py:265-286 (host meshgrid + F.grid_sample
 * of torch 0.3.1: bilinear, zero padding, ix = ((g+1)/2)*(W-1)):
 *   out[b,c,y,x] = bilinear(img[b,c], ix, iy),  ix from x+uv[b,0,y,x], iy from y+uv[b,1,y,x]
 *   img [B,C,H,W], uv [B,2,H,W] -> out [B,C,H,W]
 */
int flow_warp_forward_b200(const float *img, const float *uv, float *out,
                           int B, int C, int H, int W, void *stream);
/* g_img is accumulated with FP32 atomics and is zeroed by the callee on `stream` first: its last bits are not
 * reproducible from run to run (g_uv is).  The models never ask for g_img (the frames are network inputs); the
 * Super SloMo stages below have gather-only, deterministic adjoints. */
int flow_warp_backward_b200(const float *img, const float *uv, const float *grad_out,
                            float *g_img, float *g_uv, int B, int C, int H, int W, void *stream);

/* Replaces slomo.py:312-316: F_t0 = -(1-t)t F01 + t^2 F10; F_t1 = (1-t)^2 F01 - t(1-t) F10;
 * g0 = warp(I0, F_t0); g1 = warp(I1, F_t1).  f01,f10,f_t0,f_t1 [B,2,H,W]; i0,i1,g0,g1 [B,C,H,W]. */
int slomo_flow_combine_warp_forward_b200(const float *i0, const float *i1,
                                         const float *f01, const float *f10, double t,
                                         float *f_t0, float *f_t1, float *g0, float *g1,
                                         int B, int C, int H, int W, void *stream);

/* Replaces slomo.py:320-328: F_ref = clamp(dF+F,-1,1); V1 = 1-V0; g = warp(I,F_ref);
 * out = ((1-t)V0 g0 + t V1 g1) / ((1-t)V0 + t V1).   d_t0,d_t1 [B,2,H,W]; v_t0 [B,1,H,W]. */
int slomo_refine_blend_forward_b200(const float *i0, const float *i1,
                                    const float *f_t0, const float *f_t1,
                                    const float *d_t0, const float *d_t1, const float *v_t0,
                                    double t, float *out, int B, int C, int H, int W, void *stream);

/* The same two stages for ALL T middle frames of a batch in one launch each, with their adjoints
 * (slomo.py:307-340: the per-t loop of SloMo.forward; the time steps are independent).  t = (t_+1)/(T+1) in
 * double as in the reference (slomo.py:2,312); T <= 16.  Sample order of the T*B batch: n = t_*B + b.
 *
 * slomo_interp_input_*: slomo.py:312-318.  interp_input [T*B, 4C+4, H, W] is the refinement network's input
 * cat(I0, g(I0,F_t0), F_t0, F_t1, g(I1,F_t1), I1) (slomo.py:318), written in place; f_t{0,1}_collector
 * [B,T,2,H,W] are the model's F_t_*_collector outputs in the reference's REVERSED time order
 * (slot T-1-t_, slomo.py:332-340).  Backward: g_f01 / g_f10 [B,2,H,W] from the gradients of all three outputs
 * (collector gradients may be NULL); gather-only, no atomics, deterministic.  Gradients w.r.t. i0 / i1 are not
 * produced (network inputs): use flow_warp_backward_b200 for those. */
int slomo_interp_input_forward_b200(const float *i0, const float *i1, const float *f01, const float *f10,
                                    float *interp_input, float *f_t0_collector, float *f_t1_collector,
                                    int B, int T, int C, int H, int W, void *stream);
int slomo_interp_input_backward_b200(const float *i0, const float *i1, const float *f01, const float *f10,
                                     const float *g_interp_input, const float *g_f_t0_collector,
                                     const float *g_f_t1_collector, float *g_f01, float *g_f10,
                                     int B, int T, int C, int H, int W, void *stream);
/* slomo_refine_blend_batched_*: slomo.py:320-328 for every (t_, b).  d_t0, d_t1 [T*B,2,H,W], v_t0 [T*B,1,H,W]
 * (the refinement network's outputs, sample order n = t_*B + b); the flows are read from the collectors; pred
 * [B,T,C,H,W] in the reference's reversed time order.  Backward: gradients w.r.t. the collectors, d_t0, d_t1
 * (torch.clamp's mask: gradient passes on [-1, 1] inclusive) and v_t0; gather-only. */
int slomo_refine_blend_batched_forward_b200(const float *i0, const float *i1, const float *f_t0_collector,
                                            const float *f_t1_collector, const float *d_t0, const float *d_t1,
                                            const float *v_t0, float *pred,
                                            int B, int T, int C, int H, int W, void *stream);
int slomo_refine_blend_batched_backward_b200(const float *i0, const float *i1, const float *f_t0_collector,
                                             const float *f_t1_collector, const float *d_t0, const float *d_t1,
                                             const float *v_t0, const float *g_pred,
                                             float *g_f_t0_collector, float *g_f_t1_collector, float *g_d_t0,
                                             float *g_d_t1, float *g_v_t0,
                                             int B, int T, int C, int H, int W, void *stream);

/* ------------------------------------------------------------------------------------------
 * Decoder resampling.  N = B*C planes of H x W; results are 2H x 2W.
 *
 * Bilinear x2 upsample with the torch-0.3.1 mapping (today's align_corners=True):
 * replaces nn.Upsample(scale_factor=2, mode='bilinear') of src/models/tai/tai.py:283,337,343 and
 * src/models/slomo/slomo.py:113-149.  src = d*(in-1)/(out-1) in FP32, i0 = (int)src, i1 = i0 + (i0 < in-1).
 * The backward entry is its adjoint, evaluated as a gather (no atomics, deterministic).
 */
int upsample_bilinear2x_forward_b200(const float *in, float *out, long long N, int H, int W, void *stream);
int upsample_bilinear2x_backward_b200(const float *grad_out, float *grad_in, long long N, int H, int W, void *stream);

/* Zero-insertion unpooling fused with the residual add of DecCnn:
 * replaces  fixed_unpooling(x) + res   src/models/mcnet/mcnet.py:234-236 (the add), 240-256 (two cats,
 * clone().zero_(), two permutes):  out[n,2y,2x] = x[n,y,x] + res[n,2y,2x], out = res elsewhere.
 * x [N,H,W], res / out [N,2H,2W].  Adjoint: grad_res = grad_out (no kernel), grad_x[n,y,x] = grad_out[n,2y,2x]. */
int unpool_add_forward_b200(const float *x, const float *res, float *out, long long N, int H, int W, void *stream);
int unpool_backward_b200(const float *grad_out, float *grad_x, long long N, int H, int W, void *stream);

/* 2 x 2 max pooling, stride 2, floor mode:
 * replaces nn.MaxPool2d(2) of src/models/mcnet/mcnet.py:28-45 (motion / content encoders) and
 * src/models/slomo/slomo.py:47-85.  in [N,H,W] -> out [N,H/2,W/2] and code [N,H/2,W/2] (one byte per output:
 * 2*dy+dx of the selected element; scan order (0,0),(0,1),(1,0),(1,1), a later element wins only if greater
 * or NaN -- the library's rule, so ties behind a ReLU go to the first position).  Backward writes EVERY element
 * of grad_in [N,H,W] once (zeros for non-selected positions and for the unpooled last row / column of odd
 * sizes): the caller's buffer needs no zero-fill.  H, W >= 2. */
int maxpool2x2_forward_b200(const float *in, float *out, unsigned char *code, long long N, int H, int W, void *stream);
int maxpool2x2_backward_b200(const float *grad_out, const unsigned char *code, float *grad_in, long long N, int H, int W,
                             void *stream);

/* ------------------------------------------------------------------------------------------
 * Reconstruction losses of the training step: MSELoss + GDL of a prediction
 * against the ground truth in one pass, with the inverse transform v01 = (v + add) * mul folded in.
 * replaces, per prediction tensor, src/environments/environments.py:363-371,447-451 (permute + contiguous +
 * inverse_transform twice, torch.nn.MSELoss, GDL) and src/losses/losses.py:24-45 (sliced differences, two
 * L1Loss(reduce=False), sliced copies, add, mean).  pred / target [planes,H,W] contiguous (any plane order:
 * both results are means).  out2[0] = mean((x01-y01)^2), out2[1] = the GDL mean over planes*(H-1)*(W-1)
 * (NaN when H == 1 or W == 1, like the mean of an empty tensor).  `workspace`: device scratch of
 * l2_gdl_loss_workspace_bytes() bytes (per-CTA partial sums; the final sum is taken in a fixed order in
 * double, so the result is deterministic).
 * Backward: grad_pred = grad_mse * d mse/d pred + grad_gdl * d gdl/d pred, where grad_mse / grad_gdl are
 * DEVICE scalars (the upstream gradients of the two means; NULL = 0), read on the device: no host sync.
 * sign(0) = 0 as in torch's abs backward. */
long long l2_gdl_loss_workspace_bytes(long long planes, int H, int W);
int l2_gdl_loss_forward_b200(const float *pred, const float *target, long long planes, int H, int W, float add, float mul,
                             float *out2, void *workspace, void *stream);
int l2_gdl_loss_backward_b200(const float *pred, const float *target, long long planes, int H, int W, float add, float mul,
                              const float *grad_mse, const float *grad_gdl, float *grad_pred, void *stream);

/* ------------------------------------------------------------------------------------------
 * Bias + activation epilogue of the convolution layers and its adjoint (one pass each way).
 * replaces, around every `nn.Conv2d(bias=True)` [+ nn.ReLU / nn.LeakyReLU] of src/models/mcnet/mcnet.py:28-45,79-104,
 * 137-225, src/models/tai/tai.py:244-347 and src/models/slomo/slomo.py:28-260, the library's broadcast add_(bias),
 * clamp_min / leaky_relu, threshold_backward / leaky_relu_backward and the sum over (N,H,W) for the bias gradient.
 * act: 0 none, 1 relu, 2 leaky relu (negative slope alpha).
 * forward, IN PLACE on the convolution output y [N,C,HW]:  y = act(y + bias[c])  (same two FP32 ops as the library).
 * backward: grad_in = grad_out * act'(out) (act' from the forward OUTPUT, as the library does; grad_in may be NULL for
 * act == 0, where it would equal grad_out) and grad_bias[c] = sum over n, hw of grad_in (fixed-order two-stage sum in
 * `workspace` of bias_act_backward_workspace_bytes(N, C) bytes: deterministic). */
int bias_act_forward_b200(float *y, const float *bias, long long N, int C, int HW, int act, float alpha, void *stream);
long long bias_act_backward_workspace_bytes(long long N, int C);
int bias_act_backward_b200(const float *grad_out, const float *out, float *grad_in, float *grad_bias, void *workspace,
                           long long N, int C, int HW, int act, float alpha, void *stream);

/* out = v / (sqrt(sum v^2) + eps) for one vector of n floats (one launch):
 * replaces `_l2normalize` of src/discriminators/SNDiscriminator.py:5-7 (pow, sum, pow, add, div -- five launches,
 * called twice per power iteration of every spectral-norm layer: 1170 times per KTH training step). */
int l2_normalize_b200(const float *v, float *out, int n, float eps, void *stream);

/* ------------------------------------------------------------------------------------------
 * Gather-concatenation along batch and channel, one launch, and its adjoint.
 * Replaces the torch.cat calls that assemble the inputs of the convolution stacks (src/models/tai/tai.py:182,195;
 * src/models/mcnet/mcnet.py:79,91,148) and, with the T middle frames and the two MC-Net streams batched, the two
 * copy levels (stack over t, then cat over streams) they would need.  dst [N, dst_channels, H, W] is described by
 * blocks (at most 96 per call; the array is read on the HOST at call time):
 *     dst[dst_sample + b*dst_sample_stride, dst_channel : dst_channel+channels] <- src[src_sample + b]   b < samples
 * src == NULL fills the slot with fill_value.  A source may be a channel slice of a larger tensor (src points at
 * its first element, src_sample_stride is the parent's sample pitch).  channels*H*W must be a multiple of 4 and all pointers 16-byte
 * aligned (TAI_ERR_UNSUPPORTED otherwise).  The blocks must not overlap in dst; dst elements no block covers are
 * left untouched.  Backward: the same blocks with `src` pointing at the GRADIENT buffers of the sources, which are
 * overwritten from grad_dst (a source sample belongs to one block, so nothing needs zero-filling or atomics). */
typedef struct tai_cat_block {
    const float *src;
    long long src_sample;
    long long src_sample_stride; /* floats between source samples; 0 = dense (channels*H*W); larger for a channel slice */
    long long dst_sample;
    long long dst_sample_stride;
    long long dst_channel;
    int channels;
    int samples;
    float fill_value;
} tai_cat_block;
int gather_concat_forward_b200(const tai_cat_block *blocks, int nblocks, float *dst, long long dst_channels,
                               int H, int W, void *stream);
int gather_concat_backward_b200(const tai_cat_block *blocks, int nblocks, const float *grad_dst, long long dst_channels,
                                int H, int W, void *stream);

/* ------------------------------------------------------------------------------------------
 * Motion-stream prologue.  Replaces the elementwise chains of
 * src/models/tai/tai.py:67-74 (inverse_transform -> bgr2gray_batched -> frame differences, and the same on the
 * time-reversed following frames), src/models/mcnet/mcnet.py:439-447 (the next motion input inside the
 * prediction loop) and src/util/util.py:22-41.  FP32 with one rounding per reference operation: bit-identical.
 *
 * gray_difference_frames: frames [B,K,C,H,W] in [-1,1] -> out [B,K-1,1,H,W];
 *   out[b,k] = gray01(frame i(k+1)) - gray01(frame i(k)), i(k) = k, or K-1-k when `reverse` (no flip copy).
 * gray_difference_pair:   a, b [N,C,H,W] -> out [N,1,H,W] = gray01(a) - gray01(b); backward writes
 *   g_a = 0.5 w_c g and / or g_b = -0.5 w_c g (either may be NULL).  C must be 1 or 3 (BGR). */
int gray_difference_frames_b200(const float *frames, float *out, int B, int K, int C, int H, int W, int reverse,
                                void *stream);
int gray_difference_pair_forward_b200(const float *a, const float *b, float *out, long long N, int C, int H, int W,
                                      void *stream);
int gray_difference_pair_backward_b200(const float *grad_out, float *g_a, float *g_b, long long N, int C, int H, int W,
                                       void *stream);

/* ------------------------------------------------------------------------------------------
 * Output side: frames in [-1, 1] -> 8-bit interleaved images.
 * replaces, per frame, predict.py:124-134 (save_video_frames: torch.clamp(video, -1, 1), to_numpy,
 * (255 * inverse_transform(frame)).astype(np.uint8), [:, :, ::-1] for colour) before the PNG encoder.
 * frames [N,C,H,W] float -> out [N,H,W,C] bytes; flip_channels != 0 reverses the channel order (BGR -> RGB).
 * Same FP32 operation order as the reference, truncation toward zero: byte-exact. */
int frames_to_uint8_b200(const float *frames, unsigned char *out, long long N, int C, int H, int W, int flip_channels,
                         void *stream);

/* ------------------------------------------------------------------------------------------
 * Measurement helper: a pure FFMA loop, used by bench.py to report the on-box FP32 FMA
 * ceiling next to the nominal 148 x 128 x 2 x f_SM.  Writes one float per thread to `sink`
 * (gridDim*blockDim floats).  flops = 2 * 8 * iters * grid * block. */
int tai_b200_ffma_probe(float *sink, int grid, int block, int iters, int packed, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TAI_B200_H */
