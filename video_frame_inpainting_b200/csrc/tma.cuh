// Minimal TMA (cp.async.bulk.tensor) + mbarrier helpers for sm_100a, and host-side tensor-map
// creation through the driver entry point (no link-time dependency on libcuda).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tai {

// ---- device ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals));
}

__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// Non-blocking probe (try_wait may suspend the thread for a hardware time-out before answering false).
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// Generic-proxy writes/reads of shared memory must be ordered against the async proxy (TMA) when a
// buffer is handed back to TMA for refill.
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *tm, uint64_t *bar,
                                            int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// Pull a box into L2 only (no shared-memory destination, nothing to wait on).
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap *tm, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(tm), "r"(c0),
                 "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tm)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

// 4-byte cp.async (LDGSTS); !valid -> the destination word is zero-filled and the source is not read.
__device__ __forceinline__ void cp_async_f32(float *dst, const float *src, bool valid)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(valid ? 4 : 0)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- host --------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled()
{
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// Tensor map over a kernel map K[B][ks][Ho][Wo] (FP32) whose boxes are {bw columns, bh rows, btaps taps}
// of one sample: shared-memory image [tap][row][col].  Returns false when TMA cannot describe the
// tensor (row pitch or base not 16 B aligned, driver entry point missing).
inline bool make_kernel_map_tmap(CUtensorMap *tm, const float *base, int B, int ks, int Ho, int Wo, int bw, int bh,
                                 int btaps)
{
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    if ((Wo % 4) != 0 || (((uintptr_t)base) & 15) != 0) return false;
    cuuint64_t dims[4] = {(cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)ks, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)Wo * 4, (cuuint64_t)Ho * Wo * 4, (cuuint64_t)ks * Ho * Wo * 4};
    cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)btaps, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// The same tensor with the tap index as the SECOND box dimension and the 128-byte swizzle: shared-memory image
// [row][tap][32 cols], one 128-byte line per (row, tap), 16-byte chunk c of line rho stored at chunk c ^ (rho & 7)
// (the destination must be 1024-byte aligned).  Why: in the dense [tap][row][col] image every stride is a multiple of
// 32 banks, so the four tap groups of a warp (which read four different taps of the same 8 columns) collide 4-way when
// the taps are copied to registers; here consecutive taps get different XOR masks and the copy is 2-way.
inline bool make_kernel_map_tmap_swz(CUtensorMap *tm, const float *base, int B, int ks, int Ho, int Wo, int bh, int btaps)
{
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    if ((Wo % 4) != 0 || (((uintptr_t)base) & 15) != 0) return false;
    cuuint64_t dims[4] = {(cuuint64_t)Wo, (cuuint64_t)ks, (cuuint64_t)Ho, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)Ho * Wo * 4, (cuuint64_t)Wo * 4, (cuuint64_t)ks * Ho * Wo * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)btaps, (cuuint32_t)bh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// Per-lane offsets into a swizzled [row][tap][32] box for the lane's column `col` and tap group `ch`:
// element (row r, tap ch + 4 jj) lives at  (r * KS + 4 * jj) * 32 + ch * 32 + tbl[(r * KS + 4 * jj) & 7].
__device__ __forceinline__ void swz_table(int col, int ch, int (&tbl)[8])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) tbl[k] = ch * 32 + ((((col >> 2) ^ ((k + ch) & 7)) << 2) | (col & 3));
}

}  // namespace tai
