// Types shared by the forward separable-convolution kernels.
#pragma once

#include <type_traits>

namespace tai {

// TAI_FP / TAI_FWD_MIN_CTAS: overridable by the lab harnesses (tools/lab) only; the product builds with the defaults.
#ifndef TAI_FP
#define TAI_FP 8
#endif
#ifndef TAI_FWD_MIN_CTAS
#define TAI_FWD_MIN_CTAS 3
#endif
constexpr int FP = TAI_FP;  // output rows per thread
constexpr int FNX = 8;  // output columns per warp

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

struct FwdParams {
    const float *in[2];   // [B,C,Hi,Wi] (PAD: [B,C,Ho,Wo])
    const float *ver[2];  // [B,ks,Ho,Wo]
    const float *hor[2];
    float *out[2];        // per-stream result (DUAL: dot1/dot2, may be null)
    float *blend;         // DUAL only
    float a, b;
    int B, C, Ho, Wo, ks;
    int ntx, nty;
};

}  // namespace tai
