/*
 * TEST INFRASTRUCTURE ONLY -- definitions behind oracle/ref_shim/THC.h plus the global
 * `THCState* state` that the reference's C shim expects (SeparableConvolution_cuda.c:6).
 */
#include "THC.h"

static THCState g_state = {0};
THCState *state = &g_state;
static int g_last_error = 0;

long THCudaTensor_nElement(THCState *s, const THCudaTensor *t)
{
    (void)s;
    return t->size[0] * t->size[1] * t->size[2] * t->size[3];
}

float *THCudaTensor_data(THCState *s, const THCudaTensor *t)
{
    (void)s;
    return t->data;
}

cudaStream_t THCState_getCurrentStream(THCState *s) { return s->stream; }

void ref_shim_cuda_check(cudaError_t err, const char *file, int line)
{
    (void)file;
    (void)line;
    if (err != cudaSuccess)
        g_last_error = (int)err;
}

/* helpers for the Python harness (tests/ref_kernels.py) */
void ref_shim_set_stream(void *stream) { g_state.stream = (cudaStream_t)stream; }
int ref_shim_last_error(void)
{
    int e = g_last_error;
    g_last_error = 0;
    return e;
}
