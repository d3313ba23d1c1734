// Per-pixel separable local convolution, backward, for sm_100a.
//
//   gV[b,i,y,x] = sum_c gO[b,c,y,x] * sum_j H[b,j,y,x] * I[b,c,y+i,x+j]        (kernel.cu:49-86)
//   gH[b,j,y,x] = sum_c gO[b,c,y,x] * sum_i V[b,i,y,x] * I[b,c,y+i,x+j]        (kernel.cu:88-118)
//   gI[b,c,yy,xx] = sum_{i,j : 0<=yy-i<Ho, 0<=xx-j<Wo} gO[b,c,yy-i,xx-j] * V[b,i,yy-i,xx-j] * H[b,j,yy-i,xx-j]
//                                                                              (kernel.cu:120-162)
// The reference launches three kernels that each re-walk the ks x ks window (kernel.cu:200-239).
// Here gV and gH come out of ONE pass over the window (the same LDS of I feeds both), with the
// lane layout of the forward kernel; gI is a separate gather kernel.
#include "common.cuh"
#include "sepconv_bwd_vh_v3.cuh"
#include "sepconv_bwd_i_v4.cuh"

namespace tai {

constexpr int BNX = 8;  // output columns per warp

// ------------------------------------------------------------------------------------------------
// gV + gH.  A warp owns 8 columns x BP rows; lane = (cx = lane&7, ch = lane>>3); lane group ch owns
// taps j == ch (mod 4).  Per input row yy and output row r (vertical tap i = yy - r):
//     s_r      = sum_{j in group} H_j * I[yy][x+j]           -> gV_i needs sum over the 4 groups
//     a_r[j]  += (V_i * gO) * I[yy][x+j]                      -> gH_j, complete inside the lane
// The four partial s_r of a pixel are combined with a 3-shuffle reduce-scatter so that lane group ch
// ends up with the total for output row r = ch and stores one gV value per input row.
template <int J, int CG, int WX, int WY, bool PAD>
__global__ void __launch_bounds__(32 * WX * WY, 3)
sepconv_bwd_vh_kernel(const BwdParams p)
{
    static_assert(BP == 4, "reduce-scatter below assumes 4 rows == 4 tap groups");
    constexpr int NT = 32 * WX * WY;
    constexpr int TILE_W = WX * BNX, TILE_H = WY * BP;
    constexpr int PITCH = TILE_W + 4 * J;
    extern __shared__ __align__(1024) float smem[];

    const int ks = p.ks, Ho = p.Ho, Wo = p.Wo;
    const int Hi = Ho + ks - 1, Wi = Wo + ks - 1;
    const int rows = TILE_H + ks - 1;
    const int cstride = rows * PITCH;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cx = lane & 7, ch = lane >> 3;
    const int wx = warp % WX, wy = warp / WX;

    int t = blockIdx.x;
    const int tx = t % p.ntx;
    t /= p.ntx;
    const int ty = t % p.nty;
    const int b = t / p.nty;
    const int x0 = max(0, min(tx * TILE_W, Wo - TILE_W));
    const int y0 = min(ty * TILE_H, Ho - TILE_H);
    const int px_raw = x0 + wx * BNX + cx;
    const bool px_ok = px_raw < Wo;
    const int px = px_ok ? px_raw : Wo - 1;
    const int py0 = y0 + wy * BP;
    const long plane = (long)Ho * Wo;

    // ---- stage the input halo of all CG (== C) channels ----
    for (int c = 0; c < CG; ++c) {
        const float *src = PAD ? p.in + ((long)(b * CG + c)) * plane : p.in + ((long)(b * CG + c)) * Hi * Wi;
        for (int ry = warp; ry < rows; ry += NT / 32) {
            const int gy = y0 + ry;
            for (int rx = lane; rx < PITCH; rx += 32) {
                const int gx = x0 + rx;
                float val = 0.f;
                if (rx < TILE_W + ks - 1) {
                    if (PAD) {
                        const int sy = clampi(gy - ks / 2, 0, Ho - 1);
                        const int sx = clampi(gx - ks / 2, 0, Wo - 1);
                        val = __ldg(src + (long)sy * Wo + sx);
                    } else if (gx < Wi) {
                        val = __ldg(src + (long)gy * Wi + gx);
                    }
                }
                smem[c * cstride + ry * PITCH + rx] = val;
            }
        }
    }

    const long pix = ((long)b * ks * Ho + py0) * Wo + px;  // tap 0, row py0
    float h[BP][J], a[BP][J], go[CG][BP];
#pragma unroll
    for (int jj = 0; jj < J; ++jj) {
        const int j = ch + 4 * jj;
#pragma unroll
        for (int r = 0; r < BP; ++r) {
            h[r][jj] = (j < ks) ? ld_stream(p.hor + pix + (long)j * plane + (long)r * Wo) : 0.f;
            a[r][jj] = 0.f;
        }
    }
#pragma unroll
    for (int c = 0; c < CG; ++c)
#pragma unroll
        for (int r = 0; r < BP; ++r)
            go[c][r] = __ldg(p.gout + ((long)(b * CG + c) * Ho + py0 + r) * Wo + px);
    __syncthreads();

    const float *srow = smem + (wy * BP) * PITCH + wx * BNX + cx + ch;
    const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;

#pragma unroll 1
    for (int yy = 0; yy < ks + BP - 1; ++yy) {
        float v[BP], tsum[BP];
#pragma unroll
        for (int r = 0; r < BP; ++r) {
            const int i = yy - r;
            v[r] = (i >= 0 && i < ks) ? ld_stream(p.ver + pix + (long)i * plane + (long)r * Wo) : 0.f;
            tsum[r] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < CG; ++c) {
            float iv[J];
#pragma unroll
            for (int jj = 0; jj < J; ++jj) iv[jj] = srow[c * cstride + yy * PITCH + 4 * jj];
#pragma unroll
            for (int r = 0; r < BP; ++r) {
                const float w = v[r] * go[c][r];
                float s = h[r][0] * iv[0];
                a[r][0] = fmaf(w, iv[0], a[r][0]);
#pragma unroll
                for (int jj = 1; jj < J; ++jj) {
                    s = fmaf(h[r][jj], iv[jj], s);
                    a[r][jj] = fmaf(w, iv[jj], a[r][jj]);
                }
                tsum[r] = fmaf(go[c][r], s, tsum[r]);
            }
        }
        // reduce-scatter over the four tap groups: group ch ends with the total of row r = ch
        const float keep0 = hi16 ? tsum[2] : tsum[0];
        const float keep1 = hi16 ? tsum[3] : tsum[1];
        const float send0 = hi16 ? tsum[0] : tsum[2];
        const float send1 = hi16 ? tsum[1] : tsum[3];
        const float u0 = keep0 + __shfl_xor_sync(0xffffffffu, send0, 16);
        const float u1 = keep1 + __shfl_xor_sync(0xffffffffu, send1, 16);
        const float tot = (hi8 ? u1 : u0) + __shfl_xor_sync(0xffffffffu, hi8 ? u0 : u1, 8);
        const int i = yy - ch;
        if (p.gver && px_ok && i >= 0 && i < ks)
            p.gver[pix + (long)i * plane + (long)ch * Wo] = tot;
    }

    if (p.ghor && px_ok) {
#pragma unroll
        for (int jj = 0; jj < J; ++jj) {
            const int j = ch + 4 * jj;
            if (j < ks) {
#pragma unroll
                for (int r = 0; r < BP; ++r) p.ghor[pix + (long)j * plane + (long)r * Wo] = a[r][jj];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// gI (padded input gradient), gather form.  A thread owns GP consecutive rows of one padded column
// xx and walks the source pixels (Y, X = xx - j) that can reach them; for each it forms
// w_c = gO_c * H_j once and spends one V load per FMA row.  Lanes run along x, so every load is
// coalesced; the bounds test X<0 || Y<0 || Y>=Ho || X>=Wo of kernel.cu:150 becomes the loop limits.
constexpr int GP = 8;

template <int CG>
__global__ void __launch_bounds__(128)
sepconv_bwd_i_kernel(const BwdParams p)
{
    const int ks = p.ks, Ho = p.Ho, Wo = p.Wo;
    const int Hi = Ho + ks - 1, Wi = Wo + ks - 1;
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    const int yy0 = blockIdx.y * GP;
    const int b = blockIdx.z;
    if (xx >= Wi) return;
    const long plane = (long)Ho * Wo;

    float acc[CG][GP];
#pragma unroll
    for (int c = 0; c < CG; ++c)
#pragma unroll
        for (int r = 0; r < GP; ++r) acc[c][r] = 0.f;

    const int Ylo = max(0, yy0 - (ks - 1)), Yhi = min(Ho - 1, yy0 + GP - 1);
    const int jlo = max(0, xx - (Wo - 1)), jhi = min(ks - 1, xx);  // 0 <= X = xx - j < Wo
    for (int Y = Ylo; Y <= Yhi; ++Y) {
        const float *vrow = p.ver + ((long)b * ks * Ho + Y) * Wo;
        const float *hrow = p.hor + ((long)b * ks * Ho + Y) * Wo;
        for (int j = jlo; j <= jhi; ++j) {
            const int X = xx - j;
            const float hj = __ldg(hrow + (long)j * plane + X);
            float w[CG];
#pragma unroll
            for (int c = 0; c < CG; ++c)
                w[c] = __ldg(p.gout + ((long)(b * CG + c) * Ho + Y) * Wo + X) * hj;
#pragma unroll
            for (int r = 0; r < GP; ++r) {
                const int i = yy0 + r - Y;
                if (i >= 0 && i < ks) {
                    const float vi = __ldg(vrow + (long)i * plane + X);
#pragma unroll
                    for (int c = 0; c < CG; ++c) acc[c][r] = fmaf(vi, w[c], acc[c][r]);
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < CG; ++c)
#pragma unroll
        for (int r = 0; r < GP; ++r)
            if (yy0 + r < Hi) p.gin[((long)(b * CG + c) * Hi + yy0 + r) * Wi + xx] = acc[c][r];
}

// ------------------------------------------------------------------------------------------------
// Shape-agnostic fallbacks: the reference's three loop nests, one thread per output element.
template <bool PAD>
__device__ __forceinline__ float load_in(const BwdParams &p, int b, int c, int y, int x)
{
    if (PAD) {
        const int sy = clampi(y - p.ks / 2, 0, p.Ho - 1), sx = clampi(x - p.ks / 2, 0, p.Wo - 1);
        return p.in[((long)(b * p.C + c) * p.Ho + sy) * p.Wo + sx];
    }
    return p.in[((long)(b * p.C + c) * (p.Ho + p.ks - 1) + y) * (p.Wo + p.ks - 1) + x];
}

template <bool PAD>
__global__ void sepconv_bwd_vh_simple_kernel(const BwdParams p)
{
    const int ks = p.ks, Ho = p.Ho, Wo = p.Wo;
    const long plane = (long)Ho * Wo;
    const long n = (long)p.B * ks * plane;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
        const int x = idx % Wo;
        const int y = (idx / Wo) % Ho;
        const int tap = (idx / plane) % ks;
        const int b = idx / (plane * ks);
        float gv = 0.f, gh = 0.f;
        for (int c = 0; c < p.C; ++c) {
            const float go = p.gout[((long)(b * p.C + c) * Ho + y) * Wo + x];
            float sv = 0.f, sh = 0.f;
            for (int f = 0; f < ks; ++f) {
                const long kf = ((long)(b * ks + f) * Ho + y) * Wo + x;
                sv = fmaf(load_in<PAD>(p, b, c, y + tap, x + f), p.hor[kf], sv);
                sh = fmaf(load_in<PAD>(p, b, c, y + f, x + tap), p.ver[kf], sh);
            }
            gv = fmaf(go, sv, gv);
            gh = fmaf(go, sh, gh);
        }
        if (p.gver) p.gver[idx] = gv;
        if (p.ghor) p.ghor[idx] = gh;
    }
}

__global__ void sepconv_bwd_i_simple_kernel(const BwdParams p)
{
    const int ks = p.ks, Ho = p.Ho, Wo = p.Wo;
    const int Hi = Ho + ks - 1, Wi = Wo + ks - 1;
    const long n = (long)p.B * p.C * Hi * Wi;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
        const int xx = idx % Wi;
        const int yy = (idx / Wi) % Hi;
        const int c = (idx / ((long)Wi * Hi)) % p.C;
        const int b = idx / ((long)Wi * Hi * p.C);
        float acc = 0.f;
        for (int fx = 0; fx < ks; ++fx)
            for (int fy = 0; fy < ks; ++fy) {
                const int X = xx - (ks - 1) + fx, Y = yy - (ks - 1) + fy;
                if (X < 0 || Y < 0 || Y >= Ho || X >= Wo) continue;
                const long k = ((long)b * ks * Ho + Y) * Wo + X;
                acc += p.gout[((long)(b * p.C + c) * Ho + Y) * Wo + X] *
                       p.ver[k + (long)(ks - 1 - fy) * Ho * Wo] * p.hor[k + (long)(ks - 1 - fx) * Ho * Wo];
            }
        p.gin[idx] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// Replication pad (tai.py:170-171) forward / adjoint, and the upstream-gradient mix of the fused op.
__global__ void reppad_fwd_kernel(const float *__restrict__ in, float *__restrict__ out, int N, int H, int W, int pad)
{
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    const long n = (long)N * Hp * Wp;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
        const int xx = idx % Wp;
        const int yy = (idx / Wp) % Hp;
        const long img = idx / ((long)Wp * Hp);
        out[idx] = in[(img * H + clampi(yy - pad, 0, H - 1)) * W + clampi(xx - pad, 0, W - 1)];
    }
}

// Adjoint: every unpadded element sums the padded elements that were copied from it (fixed order,
// deterministic -- no atomics).  A group of threads owns one unpadded row: the threads sum the source rows
// column by column (coalesced; one source row for an interior output row, pad+1 for the first / last one),
// interior columns are stored directly and the pad+1 columns of each border are combined with a shuffle
// tree.  Interior rows: one WARP per row.  The first / last row of every image: one whole CTA per row, the
// pad+1 source rows unrolled so that eight loads per thread are in flight -- with a warp per row these two
// rows (26 x 178 elements at pad = 25) were a serial chain of ~150 dependent-latency loads and set the
// kernel's duration (35 us for 4 MB; a thread per output element: 51 us).
template <int GROUP>
__device__ __forceinline__ void reppad_bwd_row(const float *__restrict__ gpad, float *__restrict__ gin, long row, int H, int W,
                                               int pad, int t, float *s_red)
{
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    const int y = (int)(row % H);
    const long img = row / H;
    const int ylo = (y == 0) ? 0 : y + pad;
    const int yhi = (y == H - 1) ? Hp - 1 : y + pad;
    const float *src = gpad + (img * Hp + ylo) * Wp;
    float *dst = gin + row * W;
    float left = 0.f, right = 0.f;
    for (int xx = t; xx < Wp; xx += GROUP) {
        float cs = 0.f;
#pragma unroll 8
        for (int yy = 0; yy <= yhi - ylo; ++yy) cs += __ldg(src + (long)yy * Wp + xx);
        const int x = xx - pad;
        if (x <= 0)
            left += cs;            // padded columns 0 .. pad feed output column 0
        else if (x >= W - 1)
            right += cs;           // padded columns W-1+pad .. Wp-1 feed output column W-1
        else
            dst[x] = cs;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        left += __shfl_xor_sync(0xffffffffu, left, o);
        right += __shfl_xor_sync(0xffffffffu, right, o);
    }
    if (GROUP > 32) {  // fixed-order sum of the warps' partials
        const int w = t >> 5;
        if ((t & 31) == 0) {
            s_red[2 * w] = left;
            s_red[2 * w + 1] = right;
        }
        __syncthreads();
        left = right = 0.f;
        for (int k = 0; k < GROUP / 32; ++k) {
            left += s_red[2 * k];
            right += s_red[2 * k + 1];
        }
    }
    if (t == 0) {
        if (W == 1) {
            dst[0] = left + right;
        } else {
            dst[0] = left;
            dst[W - 1] = right;
        }
    }
}

// grid: first the border-row CTAs (2 per image; 1 if H <= 2 rows are all border rows: handled by `nb`), then
// the interior rows, 8 per CTA
__global__ void __launch_bounds__(256)
reppad_bwd_kernel(const float *__restrict__ gpad, float *__restrict__ gin, int N, int H, int W, int pad)
{
    __shared__ float s_red[16];
    const int nb = (H >= 2) ? 2 : 1;               // border rows per image
    const long border_ctas = (long)N * nb;
    if (blockIdx.x < border_ctas) {
        const long img = blockIdx.x / nb;
        const int y = (blockIdx.x % nb == 0) ? 0 : H - 1;
        reppad_bwd_row<256>(gpad, gin, img * H + y, H, W, pad, threadIdx.x, s_red);
        return;
    }
    const int ni = H - nb;                         // interior rows per image
    const long r = (long)(blockIdx.x - border_ctas) * 8 + (threadIdx.x >> 5);
    if (ni <= 0 || r >= (long)N * ni) return;
    const long img = r / ni;
    const int y = 1 + (int)(r % ni);
    reppad_bwd_row<32>(gpad, gin, img * H + y, H, W, pad, threadIdx.x & 31, nullptr);
}

static inline unsigned reppad_bwd_grid(int N, int H)
{
    const int nb = (H >= 2) ? 2 : 1;
    const long interior = (long)N * (H - nb);
    return (unsigned)((long)N * nb + (interior + 7) / 8);
}

// gD1 = a*gP + g1, gD2 = b*gP + g2 (null inputs are zero)
__global__ void fused_grad_mix_kernel(const float *gp, const float *g1, const float *g2, float *d1, float *d2,
                                      float a, float b, long n)
{
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
        const float gpv = gp ? gp[idx] : 0.f;
        d1[idx] = a * gpv + (g1 ? g1[idx] : 0.f);
        d2[idx] = b * gpv + (g2 ? g2[idx] : 0.f);
    }
}

// Algorithmic work of the backward launches.
template <bool PAD>
static void vh_work(const BwdParams &p, double *flops, double *bytes)
{
    const double px = (double)p.B * p.Ho * p.Wo;
    const double in_el = PAD ? px * p.C : (double)p.B * p.C * (p.Ho + p.ks - 1) * (p.Wo + p.ks - 1);
    const double outs = (p.gver ? 1.0 : 0.0) + (p.ghor ? 1.0 : 0.0);
    *flops = 2.0 * 2.0 * px * p.C * p.ks * p.ks;
    *bytes = 4.0 * (px * p.C + in_el + 2.0 * px * p.ks + outs * px * p.ks);
}

static void gi_work(const BwdParams &p, double *flops, double *bytes)
{
    const double px = (double)p.B * p.Ho * p.Wo;
    *flops = 2.0 * px * p.C * p.ks * p.ks;
    *bytes = 4.0 * (px * p.C + 2.0 * px * p.ks + (double)p.B * p.C * (p.Ho + p.ks - 1) * (p.Wo + p.ks - 1));
}

static inline unsigned ew_grid(long n, int block) { return (unsigned)((n + block - 1) / block > 148L * 16 ? 148 * 16 : (n + block - 1) / block < 1 ? 1 : (n + block - 1) / block); }

template <int J, int CG, bool PAD>
static int launch_vh_tiled(const BwdParams &p0, cudaStream_t st)
{
    constexpr int WX = 4, WY = 1;
    constexpr int TILE_W = WX * BNX, TILE_H = WY * BP;
    constexpr int PITCH = TILE_W + 4 * J;
    BwdParams p = p0;
    p.ntx = ceil_div(p.Wo, TILE_W);
    p.nty = ceil_div(p.Ho, TILE_H);
    const size_t smem = (size_t)CG * (TILE_H + p.ks - 1) * PITCH * sizeof(float);
    auto kern = sepconv_bwd_vh_kernel<J, CG, WX, WY, PAD>;
    static KernelConfig kcfg;
    kcfg.get(kern, 160 * 1024, 32 * WX * WY);
    double fl, by;
    vh_work<PAD>(p, &fl, &by);
    {
        TimingScope ts("sepconv_bwd_vh", st, fl, by);
        kern<<<(unsigned)((long)p.B * p.nty * p.ntx), 32 * WX * WY, smem, st>>>(p);
    }
    note_path("bwd_vh:tiled");
    return check_launch("sepconv_bwd_vh_kernel");
}

// Persistent TMA-fed gV+gH kernel (sepconv_bwd_vh_v3.cuh); +1 = shape not TMA-describable, fall back.
template <int KS, int CG, bool PAD>
static int launch_vh_v3(const BwdParams &p0, cudaStream_t st)
{
    using Cfg = VhV3Cfg<KS>;
    BwdParams p = p0;
    VhV3Maps maps;
    static_assert(Cfg::TILE_W == 32, "the swizzled H box is one 128-byte line per (row, tap)");
    if (!make_kernel_map_tmap_swz(&maps.h, p.hor, p.B, KS, p.Ho, p.Wo, Cfg::TILE_H, KS) ||
        !make_kernel_map_tmap(&maps.v, p.ver, p.B, KS, p.Ho, p.Wo, Cfg::TILE_W, Cfg::TILE_H, Cfg::CH_TAPS))
        return 1;
    p.ntx = ceil_div(p.Wo, Cfg::TILE_W);
    p.nty = ceil_div(p.Ho, Cfg::TILE_H);
    auto kern = sepconv_bwd_vh_v3_kernel<KS, CG, PAD>;
    const size_t smem = Cfg::smem_bytes(CG);
    static KernelConfig kcfg;
    const int ctas_per_sm = kcfg.get(kern, smem, Cfg::NT);
    if (ctas_per_sm < 0) return 1;
    long ctas = (long)p.B * p.nty * p.ntx;
    const long resident = (long)sm_count() * ctas_per_sm;
    if (ctas > resident) ctas = resident;
    double fl, by;
    vh_work<PAD>(p, &fl, &by);
    {
        TimingScope ts("sepconv_bwd_vh", st, fl, by);
        kern<<<(unsigned)ctas, Cfg::NT, smem, st>>>(maps, p);
    }
    note_path("bwd_vh:v3");
    return check_launch("sepconv_bwd_vh_v3_kernel");
}

template <int KS, bool PAD>
static int launch_vh_v3_c(const BwdParams &p, cudaStream_t st)
{
    return p.C == 3 ? launch_vh_v3<KS, 3, PAD>(p, st) : launch_vh_v3<KS, 1, PAD>(p, st);
}

template <bool PAD>
static int launch_vh(const BwdParams &p, cudaStream_t st)
{
    const int ks = p.ks;
    const bool tiled = ks >= 4 && ks <= 64 && p.Ho >= BP && (p.C == 1 || p.C == 3);
    if (!tiled) {
        const long n = (long)p.B * ks * p.Ho * p.Wo;
        double fl, by;
        vh_work<PAD>(p, &fl, &by);
        {
            TimingScope ts("sepconv_bwd_vh", st, fl, by);
            sepconv_bwd_vh_simple_kernel<PAD><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(p);
        }
        note_path("bwd_vh:simple");
        return check_launch("sepconv_bwd_vh_simple_kernel");
    }
    {
        int rc = 1;
        switch (ks) {
            case 51: rc = launch_vh_v3_c<51, PAD>(p, st); break;
            case 37: rc = launch_vh_v3_c<37, PAD>(p, st); break;
            case 25: rc = launch_vh_v3_c<25, PAD>(p, st); break;
            case 13: rc = launch_vh_v3_c<13, PAD>(p, st); break;
            default: break;
        }
        if (rc <= 0) return rc;
    }
    const int j = ceil_div(ks, 4);
#define TAI_VH_CASE(JJ) \
    if (j <= JJ) return p.C == 3 ? launch_vh_tiled<JJ, 3, PAD>(p, st) : launch_vh_tiled<JJ, 1, PAD>(p, st);
    TAI_VH_CASE(4)
    TAI_VH_CASE(7)
    TAI_VH_CASE(10)
    TAI_VH_CASE(13)
    TAI_VH_CASE(16)
#undef TAI_VH_CASE
    set_error("sepconv backward: ks=%d unsupported", ks);
    return TAI_ERR_UNSUPPORTED;
}

// Persistent TMA-fed scatter kernel for gI (sepconv_bwd_i_v4.cuh); +1 = not TMA-describable, fall back.
template <int KS, bool FOLD>
static int launch_gi_v4(const BwdParams &p0, cudaStream_t st)
{
    using Cfg = GiV4Cfg<KS>;
    BwdParams p = p0;
    GiV4Maps maps;
    static_assert(Cfg::TILE_W == 32, "the swizzled H box is one 128-byte line per (row, tap)");
    if (!make_kernel_map_tmap_swz(&maps.h, p.hor, p.B, KS, p.Ho, p.Wo, Cfg::TILE_H, KS) ||
        !make_kernel_map_tmap(&maps.v, p.ver, p.B, KS, p.Ho, p.Wo, Cfg::TILE_W, Cfg::TILE_H, Cfg::CH_TAPS))
        return 1;
    p.ntx = ceil_div(p.Wo, Cfg::TILE_W);
    p.nty = ceil_div(p.Ho, Cfg::TILE_H);
    auto kern = sepconv_bwd_i_v4_kernel<KS, FOLD>;
    const size_t smem = Cfg::smem_bytes();
    static KernelConfig kcfg;
    const int ctas_per_sm = kcfg.get(kern, smem, Cfg::NT);
    if (ctas_per_sm < 0) return 1;
    const size_t gi_bytes = sizeof(float) * (size_t)p.B * p.C * (p.Ho + KS - 1) * (p.Wo + KS - 1);
    cudaError_t e = cudaMemsetAsync(p.gin, 0, gi_bytes, st);  // the kernel accumulates with red.global
    if (e != cudaSuccess) {
        set_error("sepconv_bwd_i_v4: memset: %s", cudaGetErrorString(e));
        return TAI_ERR_CUDA;
    }
    long ctas = (long)p.B * p.nty * p.ntx;
    const long resident = (long)sm_count() * ctas_per_sm;
    if (ctas > resident) ctas = resident;
    double fl, by;
    gi_work(p, &fl, &by);
    {
        TimingScope ts("sepconv_bwd_i", st, fl, by);
        kern<<<(unsigned)ctas, Cfg::NT, smem, st>>>(maps, p);
    }
    note_path("bwd_i:v4");
    return check_launch("sepconv_bwd_i_v4_kernel");
}

template <int KS>
static int launch_gi_tma(const BwdParams &p, cudaStream_t st)
{
    return p.C == 1 ? launch_gi_v4<KS, true>(p, st) : launch_gi_v4<KS, false>(p, st);
}

static int launch_gi(const BwdParams &p, cudaStream_t st)
{
    const int Hi = p.Ho + p.ks - 1, Wi = p.Wo + p.ks - 1;
    {
        int rc = 1;
        switch (p.ks) {
            case 51: rc = launch_gi_tma<51>(p, st); break;
            case 37: rc = launch_gi_tma<37>(p, st); break;
            case 25: rc = launch_gi_tma<25>(p, st); break;
            case 13: rc = launch_gi_tma<13>(p, st); break;
            default: break;
        }
        if (rc <= 0) return rc;
    }
    if ((p.C == 1 || p.C == 3) && p.B <= 65535 && ceil_div(Hi, GP) <= 65535) {
        dim3 grid(ceil_div(Wi, 128), ceil_div(Hi, GP), p.B);
        double fl, by;
        gi_work(p, &fl, &by);
        {
            TimingScope ts("sepconv_bwd_i", st, fl, by);
            if (p.C == 1)
                sepconv_bwd_i_kernel<1><<<grid, 128, 0, st>>>(p);
            else
                sepconv_bwd_i_kernel<3><<<grid, 128, 0, st>>>(p);
        }
        note_path("bwd_i:gather");
        return check_launch("sepconv_bwd_i_kernel");
    }
    const long n = (long)p.B * p.C * Hi * Wi;
    double fl, by;
    gi_work(p, &fl, &by);
    {
        TimingScope ts("sepconv_bwd_i", st, fl, by);
        sepconv_bwd_i_simple_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(p);
    }
    note_path("bwd_i:simple");
    return check_launch("sepconv_bwd_i_simple_kernel");
}

}  // namespace tai

using namespace tai;

extern "C" int SeparableConvolution_cuda_backward_b200(const float *grad_output, const float *input,
                                                       const float *vertical, const float *horizontal,
                                                       float *grad_input, float *grad_vertical,
                                                       float *grad_horizontal,
                                                       int B, int C, int Hi, int Wi, int ks, void *stream)
{
    TAI_REQUIRE(grad_output && input && vertical && horizontal, TAI_ERR_INVALID_ARGUMENT,
                "SeparableConvolution_cuda_backward_b200: null input pointer");
    TAI_REQUIRE(B > 0 && C > 0 && ks > 0 && Hi >= ks && Wi >= ks, TAI_ERR_INVALID_ARGUMENT,
                "SeparableConvolution_cuda_backward_b200: bad sizes B=%d C=%d Hi=%d Wi=%d ks=%d", B, C, Hi, Wi, ks);
    const int Ho = Hi - ks + 1, Wo = Wi - ks + 1;
    TAI_REQUIRE(fits_int31((long long)B * C * Hi * Wi) && fits_int31((long long)B * ks * Ho * Wo),
                TAI_ERR_TOO_LARGE, "SeparableConvolution_cuda_backward_b200: tensor has >= 2^31 elements");
    BwdParams p{};
    p.gout = grad_output; p.in = input; p.ver = vertical; p.hor = horizontal;
    p.gver = grad_vertical; p.ghor = grad_horizontal; p.gin = grad_input;
    p.B = B; p.C = C; p.Ho = Ho; p.Wo = Wo; p.ks = ks;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = TAI_OK;
    if (grad_vertical || grad_horizontal) rc = launch_vh<false>(p, st);
    if (rc == TAI_OK && grad_input) rc = launch_gi(p, st);
    return rc;
}

extern "C" long long tai_fused_backward_workspace_bytes(int B, int C, int H, int W, int ks)
{
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || ks <= 0) return 0;
    const long long n = (long long)B * C * H * W;
    const long long npad = (long long)B * C * (H + ks - 1) * (W + ks - 1);
    return (2 * n + npad) * (long long)sizeof(float);
}

extern "C" int tai_fused_backward_b200(const float *grad_pred, const float *grad_dot1, const float *grad_dot2,
                                       const float *pred_f, const float *pred_b,
                                       const float *v1, const float *h1, const float *v2, const float *h2,
                                       float *g_pred_f, float *g_pred_b,
                                       float *g_v1, float *g_h1, float *g_v2, float *g_h2,
                                       void *workspace,
                                       int B, int C, int H, int W, int ks, float a, float b, void *stream)
{
    TAI_REQUIRE(grad_pred || grad_dot1 || grad_dot2, TAI_ERR_INVALID_ARGUMENT,
                "tai_fused_backward_b200: all upstream gradients are null");
    TAI_REQUIRE(pred_f && pred_b && v1 && h1 && v2 && h2 && workspace, TAI_ERR_INVALID_ARGUMENT,
                "tai_fused_backward_b200: null pointer");
    TAI_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && ks > 0 && (ks & 1), TAI_ERR_INVALID_ARGUMENT,
                "tai_fused_backward_b200: bad sizes B=%d C=%d H=%d W=%d ks=%d (ks must be odd)", B, C, H, W, ks);
    TAI_REQUIRE(fits_int31((long long)B * ks * H * W) && fits_int31((long long)B * C * (H + ks) * (W + ks)),
                TAI_ERR_TOO_LARGE, "tai_fused_backward_b200: tensor has >= 2^31 elements");
    cudaStream_t st = (cudaStream_t)stream;
    const long n = (long)B * C * H * W;
    float *gd1 = (float *)workspace, *gd2 = gd1 + n, *gpad = gd2 + n;
    {
        TimingScope ts("fused_grad_mix", st, 0.0, 4.0 * 5.0 * n);
        fused_grad_mix_kernel<<<ew_grid(n, 256), 256, 0, st>>>(grad_pred, grad_dot1, grad_dot2, gd1, gd2, a, b, n);
    }
    int rc = check_launch("fused_grad_mix_kernel");
    for (int s = 0; s < 2 && rc == TAI_OK; ++s) {
        BwdParams p{};
        p.gout = s ? gd2 : gd1;
        p.in = s ? pred_b : pred_f;
        p.ver = s ? v2 : v1;
        p.hor = s ? h2 : h1;
        p.gver = s ? g_v2 : g_v1;
        p.ghor = s ? g_h2 : g_h1;
        p.gin = gpad;
        p.B = B; p.C = C; p.Ho = H; p.Wo = W; p.ks = ks;
        if (p.gver || p.ghor) rc = launch_vh<true>(p, st);
        float *gdst = s ? g_pred_b : g_pred_f;
        if (rc == TAI_OK && gdst) {
            rc = launch_gi(p, st);
            if (rc == TAI_OK) {
                {
                    TimingScope ts("reppad_bwd", st, 0.0, 4.0 * ((double)B * C * (H + ks - 1) * (W + ks - 1) + n));
                    reppad_bwd_kernel<<<reppad_bwd_grid(B * C, H), 256, 0, st>>>(gpad, gdst, B * C, H, W, ks / 2);
                }
                rc = check_launch("reppad_bwd_kernel");
            }
        }
    }
    return rc;
}

extern "C" int replication_pad_forward_b200(const float *in, float *out, int N, int H, int W, int p, void *stream)
{
    TAI_REQUIRE(in && out && N > 0 && H > 0 && W > 0 && p >= 0, TAI_ERR_INVALID_ARGUMENT,
                "replication_pad_forward_b200: bad arguments");
    const long n = (long)N * (H + 2 * p) * (W + 2 * p);
    TAI_REQUIRE(fits_int31(n), TAI_ERR_TOO_LARGE, "replication_pad_forward_b200: tensor has >= 2^31 elements");
    reppad_fwd_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, N, H, W, p);
    return check_launch("reppad_fwd_kernel");
}

extern "C" int replication_pad_backward_b200(const float *grad_out, float *grad_in, int N, int H, int W, int p, void *stream)
{
    TAI_REQUIRE(grad_out && grad_in && N > 0 && H > 0 && W > 0 && p >= 0, TAI_ERR_INVALID_ARGUMENT,
                "replication_pad_backward_b200: bad arguments");
    const long n = (long)N * H * W;
    TAI_REQUIRE(fits_int31((long)N * (H + 2 * p) * (W + 2 * p)), TAI_ERR_TOO_LARGE,
                "replication_pad_backward_b200: tensor has >= 2^31 elements");
    reppad_bwd_kernel<<<reppad_bwd_grid(N, H), 256, 0, (cudaStream_t)stream>>>(grad_out, grad_in, N, H, W, p);
    return check_launch("reppad_bwd_kernel");
}
