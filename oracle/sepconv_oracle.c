/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the per-pixel separable local convolution.
 *
 * This file restates, in plain C, the arithmetic of the four CUDA kernels of the reference
 * (MichiganCOG/video-frame-inpainting):
 *
 *   forward      src/separable_convolution/cfile/SeparableConvolution_kernel.cu:19-47
 *   grad V       src/separable_convolution/cfile/SeparableConvolution_kernel.cu:49-86
 *   grad H       src/separable_convolution/cfile/SeparableConvolution_kernel.cu:88-118
 *   grad I       src/separable_convolution/cfile/SeparableConvolution_kernel.cu:120-162
 *
 * Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
 * bench.py may load it.  The product path (video_frame_inpainting_b200) never does.
 *
 * Two flavours of every function:
 *   *_f64  float inputs, every product and sum carried in double -> the correctness oracle
 *          (FP32 results are compared with it at 1e-4 relative, see tests/).
 *   *_f32  float everything, same loop nest and same product order as the reference thread
 *          body ((I*V)*H, i outer / j inner; fx outer / fy inner for grad I) -> the "port"
 *          that is timed as the CPU baseline.
 *
 * Parity status: the reference ships no tests, golden vectors or CPU path for this operator
 * (SURVEY.md section 4), so this oracle is pinned in two other ways: (1) on a GPU box against
 * the reference's own unmodified kernels compiled into oracle/_ref (tests/test_ref_kernels_gpu.py);
 * (2) against fixtures in tests/golden/ that were produced by those reference kernels on a B200.
 *
 * Layout: all tensors contiguous NCHW.
 *   input  [B, C, Hi, Wi]      vertical, horizontal [B, ks, Ho, Wo]     output [B, C, Ho, Wo]
 *   Ho = Hi - ks + 1, Wo = Wi - ks + 1   (SeparableConvolution.py:27-28)
 */
#include <stddef.h>
#include <stdint.h>

#define IN_AT(p, b, c, y, x) ((p)[(((size_t)(b) * C + (c)) * Hi + (y)) * Wi + (x)])
#define K_AT(p, b, t, y, x) ((p)[(((size_t)(b) * ks + (t)) * Ho + (y)) * Wo + (x)])
#define OUT_AT(p, b, c, y, x) ((p)[(((size_t)(b) * C + (c)) * Ho + (y)) * Wo + (x)])

/* ------------------------------------------------------------------ forward (kernel.cu:19-47) */

void oracle_sepconv_forward_f64(const float *in, const float *ver, const float *hor, double *out,
                                int B, int C, int Hi, int Wi, int ks)
{
    const int Ho = Hi - ks + 1, Wo = Wi - ks + 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < Ho; ++y)
            for (int c = 0; c < C; ++c)
                for (int x = 0; x < Wo; ++x) {
                    double acc = 0.0;
                    for (int fy = 0; fy < ks; ++fy)
                        for (int fx = 0; fx < ks; ++fx)
                            acc += (double)IN_AT(in, b, c, y + fy, x + fx) *
                                   (double)K_AT(ver, b, fy, y, x) * (double)K_AT(hor, b, fx, y, x);
                    OUT_AT(out, b, c, y, x) = acc;
                }
}

void oracle_sepconv_forward_f32(const float *in, const float *ver, const float *hor, float *out,
                                int B, int C, int Hi, int Wi, int ks)
{
    const int Ho = Hi - ks + 1, Wo = Wi - ks + 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < Ho; ++y)
            for (int c = 0; c < C; ++c)
                for (int x = 0; x < Wo; ++x) {
                    float acc = 0.0f;
                    for (int fy = 0; fy < ks; ++fy)
                        for (int fx = 0; fx < ks; ++fx)
                            acc += IN_AT(in, b, c, y + fy, x + fx) * K_AT(ver, b, fy, y, x) *
                                   K_AT(hor, b, fx, y, x);
                    OUT_AT(out, b, c, y, x) = acc;
                }
}

/* ------------------------------------------------------------------ grad V (kernel.cu:49-86)
 * gV[b,t,y,x] = sum_c sum_f gO[b,c,y,x] * I[b,c,y+t,x+f] * H[b,f,y,x]
 * (the reference names the tap index t "intDepth", kernel.cu:66,81)                            */

void oracle_sepconv_grad_vertical_f64(const float *gout, const float *in, const float *hor,
                                      double *gver, int B, int C, int Hi, int Wi, int ks)
{
    const int Ho = Hi - ks + 1, Wo = Wi - ks + 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < ks; ++t)
            for (int y = 0; y < Ho; ++y)
                for (int x = 0; x < Wo; ++x) {
                    double acc = 0.0;
                    for (int c = 0; c < C; ++c)
                        for (int f = 0; f < ks; ++f)
                            acc += (double)OUT_AT(gout, b, c, y, x) *
                                   (double)IN_AT(in, b, c, y + t, x + f) *
                                   (double)K_AT(hor, b, f, y, x);
                    K_AT(gver, b, t, y, x) = acc;
                }
}

void oracle_sepconv_grad_vertical_f32(const float *gout, const float *in, const float *hor,
                                      float *gver, int B, int C, int Hi, int Wi, int ks)
{
    const int Ho = Hi - ks + 1, Wo = Wi - ks + 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < ks; ++t)
            for (int y = 0; y < Ho; ++y)
                for (int x = 0; x < Wo; ++x) {
                    float acc = 0.0f;
                    for (int c = 0; c < C; ++c)
                        for (int f = 0; f < ks; ++f)
                            acc += OUT_AT(gout, b, c, y, x) * IN_AT(in, b, c, y + t, x + f) *
                                   K_AT(hor, b, f, y, x);
                    K_AT(gver, b, t, y, x) = acc;
                }
}

/* ------------------------------------------------------------------ grad H (kernel.cu:88-118)
 * gH[b,t,y,x] = sum_c sum_f gO[b,c,y,x] * I[b,c,y+f,x+t] * V[b,f,y,x]                          */

void oracle_sepconv_grad_horizontal_f64(const float *gout, const float *in, const float *ver,
                                        double *ghor, int B, int C, int Hi, int Wi, int ks)
{
    const int Ho = Hi - ks + 1, Wo = Wi - ks + 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < ks; ++t)
            for (int y = 0; y < Ho; ++y)
                for (int x = 0; x < Wo; ++x) {
                    double acc = 0.0;
                    for (int c = 0; c < C; ++c)
                        for (int f = 0; f < ks; ++f)
                            acc += (double)OUT_AT(gout, b, c, y, x) *
                                   (double)IN_AT(in, b, c, y + f, x + t) *
                                   (double)K_AT(ver, b, f, y, x);
                    K_AT(ghor, b, t, y, x) = acc;
                }
}

void oracle_sepconv_grad_horizontal_f32(const float *gout, const float *in, const float *ver,
                                        float *ghor, int B, int C, int Hi, int Wi, int ks)
{
    const int Ho = Hi - ks + 1, Wo = Wi - ks + 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < ks; ++t)
            for (int y = 0; y < Ho; ++y)
                for (int x = 0; x < Wo; ++x) {
                    float acc = 0.0f;
                    for (int c = 0; c < C; ++c)
                        for (int f = 0; f < ks; ++f)
                            acc += OUT_AT(gout, b, c, y, x) * IN_AT(in, b, c, y + f, x + t) *
                                   K_AT(ver, b, f, y, x);
                    K_AT(ghor, b, t, y, x) = acc;
                }
}

/* ------------------------------------------------------------------ grad I (kernel.cu:120-162)
 * For every element (yy,xx) of the PADDED input:
 *   X = xx-(ks-1)+fx, Y = yy-(ks-1)+fy;  skip when X<0 || Y<0 || Y>=Ho || X>=Wo  (kernel.cu:150)
 *   gI += gO[b,c,Y,X] * V[b,ks-1-fy,Y,X] * H[b,ks-1-fx,Y,X]          fx outer, fy inner       */

void oracle_sepconv_grad_input_f64(const float *gout, const float *ver, const float *hor,
                                   double *gin, int B, int C, int Hi, int Wi, int ks)
{
    const int Ho = Hi - ks + 1, Wo = Wi - ks + 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int yy = 0; yy < Hi; ++yy)
                for (int xx = 0; xx < Wi; ++xx) {
                    double acc = 0.0;
                    for (int fx = 0; fx < ks; ++fx)
                        for (int fy = 0; fy < ks; ++fy) {
                            const int X = xx - (ks - 1) + fx;
                            const int Y = yy - (ks - 1) + fy;
                            if (X < 0 || Y < 0 || Y >= Ho || X >= Wo)
                                continue;
                            acc += (double)OUT_AT(gout, b, c, Y, X) *
                                   (double)K_AT(ver, b, (ks - 1) - fy, Y, X) *
                                   (double)K_AT(hor, b, (ks - 1) - fx, Y, X);
                        }
                    IN_AT(gin, b, c, yy, xx) = acc;
                }
}

void oracle_sepconv_grad_input_f32(const float *gout, const float *ver, const float *hor,
                                   float *gin, int B, int C, int Hi, int Wi, int ks)
{
    const int Ho = Hi - ks + 1, Wo = Wi - ks + 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int yy = 0; yy < Hi; ++yy)
                for (int xx = 0; xx < Wi; ++xx) {
                    float acc = 0.0f;
                    for (int fx = 0; fx < ks; ++fx)
                        for (int fy = 0; fy < ks; ++fy) {
                            const int X = xx - (ks - 1) + fx;
                            const int Y = yy - (ks - 1) + fy;
                            if (X < 0 || Y < 0 || Y >= Ho || X >= Wo)
                                continue;
                            acc += OUT_AT(gout, b, c, Y, X) * K_AT(ver, b, (ks - 1) - fy, Y, X) *
                                   K_AT(hor, b, (ks - 1) - fx, Y, X);
                        }
                    IN_AT(gin, b, c, yy, xx) = acc;
                }
}

/* ------------------------------------------------------------------ integer tables
 * Number of taps that pass the bounds test of kernel.cu:150 for every padded-input element.
 * With gO = V = H = 1 the FP32 grad-I kernel must reproduce this table bit-exactly.            */

void oracle_sepconv_grad_input_tapcount(int32_t *count, int Hi, int Wi, int ks)
{
    const int Ho = Hi - ks + 1, Wo = Wi - ks + 1;
    for (int yy = 0; yy < Hi; ++yy)
        for (int xx = 0; xx < Wi; ++xx) {
            int32_t n = 0;
            for (int fx = 0; fx < ks; ++fx)
                for (int fy = 0; fy < ks; ++fy) {
                    const int X = xx - (ks - 1) + fx;
                    const int Y = yy - (ks - 1) + fy;
                    if (X < 0 || Y < 0 || Y >= Ho || X >= Wo)
                        continue;
                    ++n;
                }
            count[(size_t)yy * Wi + xx] = n;
        }
}

/* Replication pad index map, models/tai/tai.py:170-171 (torch.nn.ReplicationPad2d(p)):
 * padded element (yy,xx) reads source (clamp(yy-p,0,H-1), clamp(xx-p,0,W-1)).                  */

void oracle_replication_pad_index(int32_t *src_y, int32_t *src_x, int H, int W, int p)
{
    const int Hp = H + 2 * p, Wp = W + 2 * p;
    for (int yy = 0; yy < Hp; ++yy) {
        int sy = yy - p;
        sy = sy < 0 ? 0 : (sy > H - 1 ? H - 1 : sy);
        src_y[yy] = sy;
    }
    for (int xx = 0; xx < Wp; ++xx) {
        int sx = xx - p;
        sx = sx < 0 ? 0 : (sx > W - 1 ? W - 1 : sx);
        src_x[xx] = sx;
    }
}

int oracle_abi_version(void) { return 1; }
