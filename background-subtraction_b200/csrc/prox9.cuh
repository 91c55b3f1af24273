// prox9.cuh -- closed-form pieces of the l_inf prox of a 3x3 tile (spams.proximalFlat 'group-lasso-linf',
// /root/reference/inexact_alm_lsd.py:71-79): v = u - Proj_{l1 ball(lambda)}(u) = sign(u) min(|u|, theta).
#pragma once
#include <cuda_runtime.h>

namespace bsub {

#define SS_CE(i, j) { const float hi_ = fmaxf(u[i], u[j]), lo_ = fminf(u[i], u[j]); u[i] = hi_; u[j] = lo_; }
// clip level of the l1-ball projection of 9 non-negative values (25-comparator sorting network, descending)
__device__ __forceinline__ float ss_clip_level9(const float* a_in, float z) {
    float u[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) u[i] = a_in[i];
    SS_CE(0, 3) SS_CE(1, 7) SS_CE(2, 5) SS_CE(4, 8)
    SS_CE(0, 7) SS_CE(2, 4) SS_CE(3, 8) SS_CE(5, 6)
    SS_CE(0, 2) SS_CE(1, 3) SS_CE(4, 5) SS_CE(7, 8)
    SS_CE(1, 4) SS_CE(3, 6) SS_CE(5, 7)
    SS_CE(0, 1) SS_CE(2, 4) SS_CE(3, 5) SS_CE(6, 8)
    SS_CE(2, 3) SS_CE(4, 5) SS_CE(6, 7)
    SS_CE(1, 2) SS_CE(3, 4) SS_CE(5, 6)
    const float inv[9] = {1.f, 0.5f, 1.f / 3.f, 0.25f, 0.2f, 1.f / 6.f, 1.f / 7.f, 0.125f, 1.f / 9.f};
    float cs = 0.f, theta = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        cs += u[k];
        const float t = (cs - z) * inv[k];
        if (u[k] > t) theta = t;
    }
    return theta;
}

}  // namespace bsub
