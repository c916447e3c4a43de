// Register-resident 8-point orthonormal DCT-II / DCT-III butterflies and the 4-lane
// ("quad") split transform used for the 4:2:0 chroma block.
//
//   D[u][x] = c(u)/2 * cos((2x+1) u pi / 16),  c(0) = 1/sqrt(2), else 1   (D D^T = I)
//
// This is the transform of the reference's dct_8x8 / idct_8x8 (utils/JPEG.py:185-208,
// :332-354: cos tensor x outer(alpha,alpha)/4) and of JpegBasic.dct/idct
// (noise_layers/jpeg.py:115-145: `coff`), evaluated by even/odd decomposition
// (36 FP32 ops per 8 points instead of the 64-FMA matrix form).
#pragma once
#include <cuda_runtime.h>

namespace wm {

#define WM_H1 0.49039264020161522f   // cos( pi/16)/2
#define WM_H2 0.46193976625564337f   // cos(2pi/16)/2
#define WM_H3 0.41573480615127262f   // cos(3pi/16)/2
#define WM_H4 0.35355339059327379f   // cos(4pi/16)/2 = 1/(2 sqrt 2)
#define WM_H5 0.27778511650980114f   // cos(5pi/16)/2
#define WM_H6 0.19134171618254492f   // cos(6pi/16)/2
#define WM_H7 0.09754516100806417f   // cos(7pi/16)/2

// In-place forward transform of 8 registers: v[x] -> v[u].
__device__ __forceinline__ void dct8(float& v0, float& v1, float& v2, float& v3,
                                     float& v4, float& v5, float& v6, float& v7) {
    const float s0 = v0 + v7, s1 = v1 + v6, s2 = v2 + v5, s3 = v3 + v4;
    const float d0 = v0 - v7, d1 = v1 - v6, d2 = v2 - v5, d3 = v3 - v4;
    const float t0 = s0 + s3, t1 = s1 + s2, t2 = s1 - s2, t3 = s0 - s3;
    v0 = (t0 + t1) * WM_H4;
    v4 = (t0 - t1) * WM_H4;
    v2 = fmaf(WM_H2, t3, WM_H6 * t2);
    v6 = fmaf(WM_H6, t3, -WM_H2 * t2);
    v1 = fmaf(WM_H1, d0, fmaf(WM_H3, d1, fmaf(WM_H5, d2, WM_H7 * d3)));
    v3 = fmaf(WM_H3, d0, fmaf(-WM_H7, d1, fmaf(-WM_H1, d2, -WM_H5 * d3)));
    v5 = fmaf(WM_H5, d0, fmaf(-WM_H1, d1, fmaf(WM_H7, d2, WM_H3 * d3)));
    v7 = fmaf(WM_H7, d0, fmaf(-WM_H5, d1, fmaf(WM_H3, d2, -WM_H1 * d3)));
}

// In-place inverse (transpose) transform: v[u] -> v[x].
__device__ __forceinline__ void idct8(float& v0, float& v1, float& v2, float& v3,
                                      float& v4, float& v5, float& v6, float& v7) {
    const float a0 = (v0 + v4) * WM_H4, a1 = (v0 - v4) * WM_H4;
    const float b0 = fmaf(WM_H2, v2, WM_H6 * v6);
    const float b1 = fmaf(WM_H6, v2, -WM_H2 * v6);
    const float e0 = a0 + b0, e3 = a0 - b0, e1 = a1 + b1, e2 = a1 - b1;
    const float o0 = fmaf(WM_H1, v1, fmaf(WM_H3, v3, fmaf(WM_H5, v5, WM_H7 * v7)));
    const float o1 = fmaf(WM_H3, v1, fmaf(-WM_H7, v3, fmaf(-WM_H1, v5, -WM_H5 * v7)));
    const float o2 = fmaf(WM_H5, v1, fmaf(-WM_H1, v3, fmaf(WM_H7, v5, WM_H3 * v7)));
    const float o3 = fmaf(WM_H7, v1, fmaf(-WM_H5, v3, fmaf(WM_H3, v5, -WM_H1 * v7)));
    v0 = e0 + o0; v7 = e0 - o0;
    v1 = e1 + o1; v6 = e1 - o1;
    v2 = e2 + o2; v5 = e2 - o2;
    v3 = e3 + o3; v4 = e3 - o3;
}

__device__ __forceinline__ void dct8(float (&v)[8])  { dct8(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]); }
__device__ __forceinline__ void idct8(float (&v)[8]) { idct8(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]); }

// ---------------------------------------------------------------------------------------------
// Quad transform: an 8x8 block is spread over 4 lanes as 4x4 quadrants (lane bit `xbit`
// selects the left/right half, lane bit `ybit` the top/bottom half).  One 8-point DCT along
// an axis is split by the first butterfly stage: the low-half lane receives the partner's
// values, forms the sums s_k and produces the EVEN outputs; the high-half lane forms the
// (negated, index-reversed) differences and produces the ODD outputs.  Both halves run the
// same instruction stream — a 4x4 matrix whose entries depend on the lane's parity —
// so there is no divergence:
//     u_m = own[m] + sigma * partner[3-m]            sigma = +1 (low half) / -1 (high half)
//     out_k = sum_m M[k][m] u_m      M = D[2k][m] (low)  or  -D[2k+1][3-m] (high)
// After the pass a lane holds frequencies f = 2k + parity.  The inverse pass is the
// transpose:  w_m = sum_k M[k][m] c_k ;  own[m] = w_m - sigma * partner_w[3-m].
// ---------------------------------------------------------------------------------------------
struct QuadCoef {
    float m[4][4];
    float sigma;
};

__device__ __forceinline__ float dct_entry(int u, int x) {
    // D[u][x] from the 7 half-cosines (exact table lookup, no transcendental at run time)
    const float h[8] = {WM_H4, WM_H1, WM_H2, WM_H3, WM_H4, WM_H5, WM_H6, WM_H7};
    if (u == 0) return WM_H4;
    int a = ((2 * x + 1) * u) & 31;          // angle in units of pi/16, period 32
    float sgn = 1.f;
    if (a > 16) a = 32 - a;                  // cos(2pi - t) = cos t
    if (a > 8) { a = 16 - a; sgn = -1.f; }   // cos(pi - t) = -cos t
    return a == 8 ? 0.f : sgn * (a == 0 ? 0.5f : h[a]);
}

__device__ __forceinline__ void quad_coef_init(QuadCoef& q, int parity) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int m = 0; m < 4; ++m)
            q.m[k][m] = parity ? -dct_entry(2 * k + 1, 3 - m) : dct_entry(2 * k, m);
    q.sigma = parity ? -1.f : 1.f;
}

// forward pass along the column index of p (horizontal): partner = lane ^ xor_mask
__device__ __forceinline__ void quad_dct_rows(float (&p)[4][4], const QuadCoef& q, int xor_mask) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float r[4], u[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) r[j] = __shfl_xor_sync(0xffffffffu, p[i][j], xor_mask);
#pragma unroll
        for (int m = 0; m < 4; ++m) u[m] = fmaf(q.sigma, r[3 - m], p[i][m]);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            p[i][k] = fmaf(q.m[k][0], u[0], fmaf(q.m[k][1], u[1], fmaf(q.m[k][2], u[2], q.m[k][3] * u[3])));
    }
}
__device__ __forceinline__ void quad_idct_rows(float (&p)[4][4], const QuadCoef& q, int xor_mask) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float w[4], r[4];
#pragma unroll
        for (int m = 0; m < 4; ++m)
            w[m] = fmaf(q.m[0][m], p[i][0], fmaf(q.m[1][m], p[i][1], fmaf(q.m[2][m], p[i][2], q.m[3][m] * p[i][3])));
#pragma unroll
        for (int m = 0; m < 4; ++m) r[m] = __shfl_xor_sync(0xffffffffu, w[m], xor_mask);
#pragma unroll
        for (int m = 0; m < 4; ++m) p[i][m] = fmaf(-q.sigma, r[3 - m], w[m]);
    }
}
// the same along the row index of p (vertical)
__device__ __forceinline__ void quad_dct_cols(float (&p)[4][4], const QuadCoef& q, int xor_mask) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float r[4], u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = __shfl_xor_sync(0xffffffffu, p[i][j], xor_mask);
#pragma unroll
        for (int m = 0; m < 4; ++m) u[m] = fmaf(q.sigma, r[3 - m], p[m][j]);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            p[k][j] = fmaf(q.m[k][0], u[0], fmaf(q.m[k][1], u[1], fmaf(q.m[k][2], u[2], q.m[k][3] * u[3])));
    }
}
__device__ __forceinline__ void quad_idct_cols(float (&p)[4][4], const QuadCoef& q, int xor_mask) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float w[4], r[4];
#pragma unroll
        for (int m = 0; m < 4; ++m)
            w[m] = fmaf(q.m[0][m], p[0][j], fmaf(q.m[1][m], p[1][j], fmaf(q.m[2][m], p[2][j], q.m[3][m] * p[3][j])));
#pragma unroll
        for (int m = 0; m < 4; ++m) r[m] = __shfl_xor_sync(0xffffffffu, w[m], xor_mask);
#pragma unroll
        for (int m = 0; m < 4; ++m) p[m][j] = fmaf(-q.sigma, r[3 - m], w[m]);
    }
}

}  // namespace wm
