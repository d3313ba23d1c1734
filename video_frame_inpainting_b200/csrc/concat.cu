// Gather-concatenation of NCHW tensors along batch and channel in one pass, and its adjoint, for sm_100a.
//
// The reference assembles the inputs of its convolution stacks with torch.cat (tai.py:182,195; mcnet.py:79,91,148)
// and, per middle frame, slices of per-time-step lists (tai.py:91-105).  With the two MC-Net streams run as one
// batch of 2B clips and the kernel network run once over the T*B middle frames the
// same data would be copied twice: once to stack the T per-step tensors of a stream, once more to concatenate
// the two streams along channels.  Here the destination is described as a list of BLOCKS
//     dst[n0 + b*nstride, coff : coff + C_k, :, :]  <-  src_k[soff + b, :, :, :]        b = 0..nb-1
// (src_k may be a channel slice of a larger tensor: its samples are then further apart than C_k*HW floats)
// (or a constant fill), each sample a contiguous run of C_k*HW floats, and ONE launch moves all of them with
// 128-bit accesses; the adjoint is the same launch with source and destination swapped (every source sample
// belongs to exactly one block, so the gradient of each source tensor is fully overwritten: no zero fill, no
// atomics).  Pure HBM streaming: 8 bytes per element.
#include "common.cuh"

namespace tai {

constexpr int kMaxBlocks = 96;

struct CatBlock {
    const float *src;   // nullptr: constant fill
    long dst_off;       // n0 * Ctot * HW + coff * HW   (floats)
    long dst_stride;    // nstride * Ctot * HW
    long src_off;       // soff * src_stride
    long src_stride;    // floats between consecutive source samples (C_k * HW when dense; larger for a channel slice)
    int run4;           // C_k * HW / 4   (float4 per sample)
    int nb;             // samples in this block
    float value;        // fill value when src == nullptr
};

struct CatPlan {
    CatBlock blk[kMaxBlocks];
    int n;
};

// grid.y = block index, grid.x strides over (sample, float4) of the block
template <bool BACKWARD>
__global__ void __launch_bounds__(256)
gather_concat_kernel(const __grid_constant__ CatPlan plan, float *__restrict__ dst)
{
    const CatBlock &k = plan.blk[blockIdx.y];
    const long total = (long)k.nb * k.run4;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long b = i / k.run4, e = i - b * k.run4;
        float4 *d = reinterpret_cast<float4 *>(dst + k.dst_off + b * k.dst_stride) + e;
        if (BACKWARD) {
            if (k.src) {
                float4 *s = reinterpret_cast<float4 *>(const_cast<float *>(k.src) + k.src_off + b * k.src_stride) + e;
                *s = __ldcs(d);
            }
        } else if (k.src) {
            *d = __ldcs(reinterpret_cast<const float4 *>(k.src + k.src_off + b * k.src_stride) + e);
        } else {
            *d = make_float4(k.value, k.value, k.value, k.value);
        }
    }
}

}  // namespace tai

using namespace tai;

// Host-side block description (plain C struct of the ABI, include/tai_b200.h: tai_cat_block).
static int run_plan(const char *who, const tai_cat_block *blocks, int nblocks, float *dst, long long Ctot, int H, int W,
                    bool backward, void *stream)
{
    TAI_REQUIRE(blocks && dst && nblocks > 0 && Ctot > 0 && H > 0 && W > 0, TAI_ERR_INVALID_ARGUMENT, "%s: bad arguments", who);
    TAI_REQUIRE(nblocks <= kMaxBlocks, TAI_ERR_UNSUPPORTED, "%s: %d blocks (limit %d)", who, nblocks, kMaxBlocks);
    const long hw = (long)H * W;
    TAI_REQUIRE((((uintptr_t)dst) & 15) == 0, TAI_ERR_UNSUPPORTED, "%s: destination not 16-byte aligned", who);
    CatPlan plan;
    plan.n = nblocks;
    long max_items = 1;
    double bytes = 0.0;
    for (int i = 0; i < nblocks; ++i) {
        const tai_cat_block &b = blocks[i];
        TAI_REQUIRE(b.channels > 0 && b.samples > 0 && b.src_sample >= 0 && b.dst_sample >= 0 && b.dst_channel >= 0 &&
                        b.dst_channel + b.channels <= Ctot && b.dst_sample_stride >= 0,
                    TAI_ERR_INVALID_ARGUMENT, "%s: block %d is malformed", who, i);
        const long run = (long)b.channels * hw;
        // every sample run and every offset must be a whole number of 16-byte vectors
        TAI_REQUIRE(run % 4 == 0 && (b.dst_channel * hw) % 4 == 0 && (Ctot * hw) % 4 == 0 &&
                        (((uintptr_t)b.src) & 15) == 0,
                    TAI_ERR_UNSUPPORTED, "%s: block %d is not 16-byte granular (C*H*W must be a multiple of 4)", who, i);
        TAI_REQUIRE(fits_int31(run) && fits_int31((long long)b.samples * run), TAI_ERR_TOO_LARGE, "%s: block %d too large", who, i);
        CatBlock &k = plan.blk[i];
        k.src = b.src;
        k.dst_off = ((long)b.dst_sample * Ctot + b.dst_channel) * hw;
        k.dst_stride = (long)b.dst_sample_stride * Ctot * hw;
        const long sstride = b.src_sample_stride > 0 ? (long)b.src_sample_stride : run;
        TAI_REQUIRE(sstride >= run && sstride % 4 == 0, TAI_ERR_INVALID_ARGUMENT, "%s: block %d: bad source sample stride", who, i);
        k.src_stride = sstride;
        k.src_off = (long)b.src_sample * sstride;
        k.run4 = (int)(run / 4);
        k.nb = b.samples;
        k.value = b.fill_value;
        max_items = max_items > (long)k.nb * k.run4 ? max_items : (long)k.nb * k.run4;
        bytes += (b.src ? 8.0 : 4.0) * b.samples * run;
    }
    cudaStream_t st = (cudaStream_t)stream;
    long gx = (max_items + 256L * 4 - 1) / (256L * 4);   // ~4 vectors per thread
    const long cap = (long)sm_count() * 8 / nblocks + 1;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)nblocks);
    {
        TimingScope ts(backward ? "gather_concat_bwd" : "gather_concat", st, 0.0, bytes);
        if (backward)
            gather_concat_kernel<true><<<grid, 256, 0, st>>>(plan, dst);
        else
            gather_concat_kernel<false><<<grid, 256, 0, st>>>(plan, dst);
    }
    return check_launch("gather_concat_kernel");
}

extern "C" int gather_concat_forward_b200(const tai_cat_block *blocks, int nblocks, float *dst, long long dst_channels,
                                          int H, int W, void *stream)
{
    return run_plan("gather_concat_forward_b200", blocks, nblocks, dst, dst_channels, H, W, false, stream);
}

extern "C" int gather_concat_backward_b200(const tai_cat_block *blocks, int nblocks, const float *grad_dst,
                                           long long dst_channels, int H, int W, void *stream)
{
    // blocks[i].src are the GRADIENT buffers of the sources here (written); fill blocks are skipped
    return run_plan("gather_concat_backward_b200", blocks, nblocks, const_cast<float *>(grad_dst), dst_channels, H, W, true,
                    stream);
}
