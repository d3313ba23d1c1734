// Development harness (not part of the product): builds ONE instantiation of a forward kernel in a
// few seconds, checks it against a naive double-precision GPU kernel and times it.
//   nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
//        -I video_frame_inpainting_b200/csrc tools/lab/fwd_lab.cu -o tools/lab/fwd_lab
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cmath>

#include "sepconv_fwd_v3.cuh"

namespace tai {
void set_error(const char *fmt, ...) { fprintf(stderr, "error: %s\n", fmt); }
void count_launch(int) {}
}  // namespace tai
using namespace tai;

#ifndef LAB_KS
#define LAB_KS 51
#endif
#ifndef LAB_CG
#define LAB_CG 1
#endif

__global__ void naive_fwd(const float *in, const float *ver, const float *hor, float *out, int B, int C, int Ho, int Wo, int ks)
{
    long n = (long)B * C * Ho * Wo;
    int Hi = Ho + ks - 1, Wi = Wo + ks - 1;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
        int x = idx % Wo, y = (idx / Wo) % Ho, c = (idx / ((long)Wo * Ho)) % C, b = idx / ((long)Wo * Ho * C);
        double acc = 0;
        for (int i = 0; i < ks; ++i) {
            double rs = 0;
            for (int j = 0; j < ks; ++j)
                rs += (double)in[((long)(b * C + c) * Hi + y + i) * Wi + x + j] * hor[((long)(b * ks + j) * Ho + y) * Wo + x];
            acc += rs * ver[((long)(b * ks + i) * Ho + y) * Wo + x];
        }
        out[idx] = (float)acc;
    }
}

static float frand(uint64_t &s)
{
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    return ((s >> 40) & 0xFFFFFF) / (float)0x1000000 * 2.f - 1.f;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

static int g_ctas_per_sm = 3;
static void run_variant(const char *name, FwdParams p, float *d_out, const std::vector<float> &ref, float *flush, size_t flush_bytes)
{
    constexpr int KS = LAB_KS, CG = LAB_CG;
    using Cfg = FwdV3Cfg<KS>;
    FwdV3Maps maps;
    if (!make_kernel_map_tmap_swz(&maps.h[0], p.hor[0], p.B, KS, p.Ho, p.Wo, Cfg::TILE_H, KS) ||   // swizzled [row][tap][col] H box
        !make_kernel_map_tmap(&maps.v[0], p.ver[0], p.B, KS, p.Ho, p.Wo, Cfg::TILE_W, Cfg::TILE_H, Cfg::CH_TAPS)) { printf("tensor map failed\n"); return; }
    maps.h[1] = maps.h[0]; maps.v[1] = maps.v[0];
    p.ntx = ceil_div(p.Wo, Cfg::TILE_W);
    p.nty = ceil_div(p.Ho, Cfg::TILE_H);
    auto kern = sepconv_fwd_v3_kernel<KS, CG, false, false>;
    size_t smem = Cfg::smem_bytes(CG);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, smem));
    long blocks = (long)p.B * p.nty * p.ntx;
    int per_sm = occ < g_ctas_per_sm ? occ : g_ctas_per_sm;
    if (blocks > 148L * per_sm) blocks = 148L * per_sm;
    size_t n = (size_t)p.B * p.C * p.Ho * p.Wo;
    CK(cudaMemset(d_out, 0, n * 4));
    kern<<<(unsigned)blocks, 128, smem>>>(maps, p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> out(n);
    CK(cudaMemcpy(out.data(), d_out, n * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, rms = 0;
    for (size_t i = 0; i < n; ++i) rms += (double)ref[i] * ref[i];
    rms = sqrt(rms / n);
    for (size_t i = 0; i < n; ++i) maxerr = std::max(maxerr, fabs((double)out[i] - ref[i]) / std::max((double)fabs(ref[i]), rms));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<float> ts;
    for (int it = 0; it < 12; ++it) {
        CK(cudaMemsetAsync(flush, it, flush_bytes));
        cudaEventRecord(e0);
        kern<<<(unsigned)blocks, 128, smem>>>(maps, p);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        ts.push_back(ms);
    }
#ifdef TAI_LAB_TIMING
    {
        unsigned long long z[8] = {0}, ph[8];
        cudaMemcpyToSymbol(g_lab_phase, z, sizeof(z));
        kern<<<(unsigned)blocks, 128, smem>>>(maps, p);
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(ph, g_lab_phase, sizeof(ph));
        double tiles = (double)p.B * p.nty * p.ntx;
        printf("   cycles per tile: issue %.0f | H-wait %.0f | H->reg+halo %.0f | V-wait %.0f | sweep %.0f | barrier %.0f | store %.0f\n",
               ph[0] / tiles, ph[1] / tiles, ph[2] / tiles, ph[3] / tiles, ph[4] / tiles, ph[5] / tiles, ph[6] / tiles);
    }
#endif
    std::sort(ts.begin(), ts.end());
    double flops = 2.0 * n * KS * KS;
    double peak = 148.0 * 128 * 2 * 1.965e9;
    printf("%-10s occ=%d ctas=%ld smem=%zu relerr=%.2e  med=%.4f ms best=%.4f ms  %.2f TFLOP/s  %.1f%% of nominal FMA peak\n",
           name, occ, blocks, smem, maxerr, ts[ts.size() / 2], ts[0], flops / (ts[ts.size() / 2] * 1e-3) / 1e12,
           100.0 * flops / (ts[ts.size() / 2] * 1e-3) / peak);
}

int main(int argc, char **argv)
{
    int B = argc > 1 ? atoi(argv[1]) : 32, Ho = argc > 2 ? atoi(argv[2]) : 128, Wo = argc > 3 ? atoi(argv[3]) : 128;
    const int C = LAB_CG, ks = LAB_KS;
    int Hi = Ho + ks - 1, Wi = Wo + ks - 1;
    size_t n_in = (size_t)B * C * Hi * Wi, n_k = (size_t)B * ks * Ho * Wo, n_out = (size_t)B * C * Ho * Wo;
    std::vector<float> h_in(n_in), h_v(n_k), h_h(n_k);
    uint64_t seed = 1234;
    for (auto &v : h_in) v = frand(seed);
    float sc = 1.f / sqrtf((float)ks);
    for (auto &v : h_v) v = frand(seed) * sc;
    for (auto &v : h_h) v = frand(seed) * sc;
    float *d_in, *d_v, *d_h, *d_out, *d_ref, *flush;
    size_t flush_bytes = 256u << 20;
    CK(cudaMalloc(&d_in, n_in * 4)); CK(cudaMalloc(&d_v, n_k * 4)); CK(cudaMalloc(&d_h, n_k * 4));
    CK(cudaMalloc(&d_out, n_out * 4)); CK(cudaMalloc(&d_ref, n_out * 4)); CK(cudaMalloc(&flush, flush_bytes));
    CK(cudaMemcpy(d_in, h_in.data(), n_in * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_v, h_v.data(), n_k * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_h, h_h.data(), n_k * 4, cudaMemcpyHostToDevice));
    naive_fwd<<<148 * 8, 256>>>(d_in, d_v, d_h, d_ref, B, C, Ho, Wo, ks);
    CK(cudaDeviceSynchronize());
    std::vector<float> ref(n_out);
    CK(cudaMemcpy(ref.data(), d_ref, n_out * 4, cudaMemcpyDeviceToHost));
    FwdParams p{};
    p.in[0] = d_in; p.ver[0] = d_v; p.hor[0] = d_h; p.out[0] = d_out;
    p.B = B; p.C = C; p.Ho = Ho; p.Wo = Wo; p.ks = ks;
    printf("shape B=%d C=%d %dx%d ks=%d\n", B, C, Ho, Wo, ks);
    printf("rows per warp FP=%d, launch bound %d CTAs/SM\n", FP, TAI_FWD_MIN_CTAS);
    g_ctas_per_sm = 8;   // as many as the occupancy query allows
    run_variant("v3 occ", p, d_out, ref, flush, flush_bytes);
    g_ctas_per_sm = 3;
    run_variant("v3 x3", p, d_out, ref, flush, flush_bytes);
    return 0;
}
