/*
 * TEST INFRASTRUCTURE ONLY -- minimal stand-in for the torch-0.3 THC headers.
 *
 * The reference's native code (src/separable_convolution/cfile/SeparableConvolution_kernel.cu
 * and SeparableConvolution_cuda.c) includes <THC.h> / <THCGeneral.h>, which no longer exist in
 * torch >= 1.0.  It touches exactly six THC names (kernel.cu:164-242, cuda.c:6): THCState,
 * THCudaTensor (->size[], ->stride[]), THCudaTensor_nElement, THCudaTensor_data,
 * THCState_getCurrentStream and THCudaCheck.  This header supplies them so that the reference
 * sources compile UNMODIFIED, from where they lie under /root/reference, into oracle/_ref
 * (see oracle/Makefile).  No reference source is copied into this repository.
 */
#ifndef REF_SHIM_THC_H
#define REF_SHIM_THC_H

#include <cuda_runtime.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct THCState {
    cudaStream_t stream;
} THCState;

typedef struct THCudaTensor {
    float *data;
    long size[4];
    long stride[4];
} THCudaTensor;

long THCudaTensor_nElement(THCState *state, const THCudaTensor *t);
float *THCudaTensor_data(THCState *state, const THCudaTensor *t);
cudaStream_t THCState_getCurrentStream(THCState *state);
void ref_shim_cuda_check(cudaError_t err, const char *file, int line);

#define THCudaCheck(err) ref_shim_cuda_check((err), __FILE__, __LINE__)

#ifdef __cplusplus
}
#endif

#endif
