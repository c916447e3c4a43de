// Coordinate / tap arithmetic of F.interpolate(align_corners=False) as ATen evaluates it in fp32
// (upsample_bilinear2d / upsample_bicubic2d, A = -0.75), shared by resize.cu and cropresize.cu.
#pragma once
#include "wm_common.cuh"

namespace wm {

// ATen area_pixel_compute_source_index (align_corners = false)
__device__ __forceinline__ float src_coord(float scale, int o) { return scale * (o + 0.5f) - 0.5f; }
__device__ __forceinline__ float cubic1(float x) { const float A = -0.75f; return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x) { const float A = -0.75f; return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

// taps of output index o along an axis of n_in samples (MODE 0: 2 taps, MODE 1: 4 taps)
template <int MODE>
__device__ __forceinline__ void taps(float scale, int o, int n_in, int (&idx)[4], float (&w)[4]) {
    float rho = src_coord(scale, o);
    if (MODE == 0) {
        rho = fmaxf(rho, 0.f);
        const int i0 = min(int(rho), n_in - 1);
        const int i1 = min(i0 + 1, n_in - 1);
        const float l1 = fminf(fmaxf(rho - i0, 0.f), 1.f);
        idx[0] = i0; idx[1] = i1; idx[2] = i1; idx[3] = i1;
        w[0] = 1.f - l1; w[1] = l1; w[2] = 0.f; w[3] = 0.f;
    } else {
        const float fl = floorf(rho);
        const int i0 = int(fl);
        const float t = rho - fl;
        w[0] = cubic2(t + 1.f); w[1] = cubic1(t); w[2] = cubic1(1.f - t); w[3] = cubic2(2.f - t);
#pragma unroll
        for (int k = 0; k < 4; ++k) idx[k] = min(max(i0 - 1 + k, 0), n_in - 1);
    }
}

// total weight with which input sample `i` enters output `o`
template <int MODE>
__device__ __forceinline__ float weight_of(float scale, int o, int n_in, int i) {
    int idx[4]; float w[4];
    taps<MODE>(scale, o, n_in, idx, w);
    float s = 0.f;
    constexpr int NT = MODE == 0 ? 2 : 4;
#pragma unroll
    for (int k = 0; k < NT; ++k) s += (idx[k] == i) ? w[k] : 0.f;
    return s;
}

// conservative candidate range of outputs whose taps can touch input i
template <int MODE>
__device__ __forceinline__ void cand_range(float inv_scale, int i, int n_in, int n_out, int& lo, int& hi) {
    const float reach = MODE == 0 ? 1.f : 2.f;
    lo = (i == 0) ? 0 : max(0, int(floorf((i - reach + 0.5f) * inv_scale - 0.5f)) - 1);
    hi = (i == n_in - 1) ? n_out - 1 : min(n_out - 1, int(ceilf((i + reach + 0.5f) * inv_scale - 0.5f)) + 1);
}

}  // namespace wm
