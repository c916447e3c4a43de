// Bilinear / bicubic interpolation with F.interpolate(size=..., align_corners=False) semantics
// (ATen upsample_bilinear2d / upsample_bicubic2d, A = -0.75), forward and exact transpose.
//
// Replaces: Resize.forward (noise_layers/resize.py:38-53: two F.interpolate + clamp) and
// Crop.forward (noise_layers/crop.py:48-53: slice + bilinear F.interpolate).  The source window
// arguments let Crop read the crop rectangle in place; `clamp01` fuses Resize's clamp.
// The backward is a gather (each input pixel sums the outputs whose taps touch it, weights
// recomputed with the SAME fp32 coordinate arithmetic as the forward), so it is deterministic —
// ATen's upsample backward uses atomicAdd.
#include "wm_common.cuh"

namespace wm {

// ATen area_pixel_compute_source_index (align_corners = false)
__device__ __forceinline__ float src_coord(float scale, int o) { return scale * (o + 0.5f) - 0.5f; }

__device__ __forceinline__ float cubic1(float x) { const float A = -0.75f; return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x) { const float A = -0.75f; return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

// taps of output index o along an axis of n_in samples: idx[0..NT), w[0..NT)
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

struct InterpArgs {
    const float* x; int64_t x_sp, x_sh; int h0, w0, Hin, Win;
    float* y; int N, Hout, Wout; float sh, sw; int clamp01;
};

template <int MODE>
__global__ void __launch_bounds__(256) interp_fwd_kernel(const InterpArgs a) {
    const int64_t total = int64_t(a.N) * a.Hout * a.Wout;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int ox = int(i % a.Wout), oy = int((i / a.Wout) % a.Hout), n = int(i / (int64_t(a.Wout) * a.Hout));
        int iy[4], ix[4]; float wy[4], wx[4];
        taps<MODE>(a.sh, oy, a.Hin, iy, wy);
        taps<MODE>(a.sw, ox, a.Win, ix, wx);
        const float* src = a.x + int64_t(n) * a.x_sp + int64_t(a.h0) * a.x_sh + a.w0;
        float acc = 0.f;
        constexpr int NT = MODE == 0 ? 2 : 4;
#pragma unroll
        for (int r = 0; r < NT; ++r) {
            const float* row = src + int64_t(iy[r]) * a.x_sh;
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < NT; ++c) s = fmaf(wx[c], __ldg(row + ix[c]), s);
            acc = fmaf(wy[r], s, acc);
        }
        if (a.clamp01) acc = fminf(fmaxf(acc, 0.f), 1.f);
        a.y[i] = acc;
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

struct InterpBwdArgs {
    const float* gy; const float* pre; int N, Hout, Wout;
    float* gx; int Hsrc, Wsrc, h0, w0, Hin, Win;
    float sh, sw, ish, isw; float* ws;
};

// pass 1: ws[n][iy][ox] = sum_oy Wy(oy, iy) * g[n][oy][ox]
template <int MODE>
__global__ void __launch_bounds__(256) interp_bwd_rows_kernel(const InterpBwdArgs a) {
    const int64_t total = int64_t(a.N) * a.Hin * a.Wout;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int ox = int(i % a.Wout), iy = int((i / a.Wout) % a.Hin), n = int(i / (int64_t(a.Wout) * a.Hin));
        int lo, hi;
        cand_range<MODE>(a.ish, iy, a.Hin, a.Hout, lo, hi);
        const int64_t base = int64_t(n) * a.Hout * a.Wout + ox;
        float acc = 0.f;
        for (int oy = lo; oy <= hi; ++oy) {
            const float w = weight_of<MODE>(a.sh, oy, a.Hin, iy);
            if (w != 0.f) {
                float g = __ldg(a.gy + base + int64_t(oy) * a.Wout);
                if (a.pre) { const float p = __ldg(a.pre + base + int64_t(oy) * a.Wout); g = (p >= 0.f && p <= 1.f) ? g : 0.f; }
                acc = fmaf(w, g, acc);
            }
        }
        a.ws[i] = acc;
    }
}

// pass 2: gx[n][h0+iy][w0+ix] = sum_ox Wx(ox, ix) * ws[n][iy][ox]; zero outside the window
template <int MODE>
__global__ void __launch_bounds__(256) interp_bwd_cols_kernel(const InterpBwdArgs a) {
    const int64_t total = int64_t(a.N) * a.Hsrc * a.Wsrc;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int x = int(i % a.Wsrc), y = int((i / a.Wsrc) % a.Hsrc), n = int(i / (int64_t(a.Wsrc) * a.Hsrc));
        const int ix = x - a.w0, iy = y - a.h0;
        float acc = 0.f;
        if (ix >= 0 && ix < a.Win && iy >= 0 && iy < a.Hin) {
            int lo, hi;
            cand_range<MODE>(a.isw, ix, a.Win, a.Wout, lo, hi);
            const float* row = a.ws + (int64_t(n) * a.Hin + iy) * a.Wout;
            for (int ox = lo; ox <= hi; ++ox) {
                const float w = weight_of<MODE>(a.sw, ox, a.Win, ix);
                if (w != 0.f) acc = fmaf(w, row[ox], acc);
            }
        }
        a.gx[i] = acc;
    }
}

static inline unsigned rs_grid(int64_t total) {
    const int64_t want = (total + 255) / 256, cap = int64_t(sm_count()) * 32;
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace wm

using namespace wm;

extern "C" int wm_interp_fwd(const float* x, int64_t x_sp, int64_t x_sh, int h0, int w0, int Hin, int Win,
                             float* y, int N, int Hout, int Wout, int mode, int clamp01, void* stream) {
    WM_REQUIRE(x && y, WM_E_NULL, "wm_interp_fwd: null pointer");
    WM_REQUIRE(mode == 0 || mode == 1, WM_E_ARG, "wm_interp_fwd: mode must be 0 (bilinear) or 1 (bicubic)");
    WM_REQUIRE(N >= 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0 && h0 >= 0 && w0 >= 0, WM_E_SHAPE,
               "wm_interp_fwd: bad shape N=%d in=%dx%d out=%dx%d", N, Hin, Win, Hout, Wout);
    if (N == 0) return WM_OK;
    InterpArgs a{x, x_sp, x_sh, h0, w0, Hin, Win, y, N, Hout, Wout,
                 (float)Hin / (float)Hout, (float)Win / (float)Wout, clamp01};
    const unsigned grid = rs_grid(int64_t(N) * Hout * Wout);
    if (mode == 0) interp_fwd_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else interp_fwd_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    WM_LAUNCH_CHECK("wm_interp_fwd");
    return WM_OK;
}

extern "C" int wm_interp_bwd(const float* gy, const float* pre, int N, int Hout, int Wout,
                             float* gx, int Hsrc, int Wsrc, int h0, int w0, int Hin, int Win,
                             int mode, float* workspace, void* stream) {
    WM_REQUIRE(gy && gx && workspace, WM_E_NULL, "wm_interp_bwd: null pointer (workspace of N*Hin*Wout floats is required)");
    WM_REQUIRE(mode == 0 || mode == 1, WM_E_ARG, "wm_interp_bwd: mode must be 0 (bilinear) or 1 (bicubic)");
    WM_REQUIRE(N >= 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0 && h0 >= 0 && w0 >= 0 &&
               h0 + Hin <= Hsrc && w0 + Win <= Wsrc, WM_E_SHAPE,
               "wm_interp_bwd: bad shape N=%d window=%dx%d@(%d,%d) src=%dx%d out=%dx%d", N, Hin, Win, h0, w0, Hsrc, Wsrc, Hout, Wout);
    if (N == 0) return WM_OK;
    InterpBwdArgs a{gy, pre, N, Hout, Wout, gx, Hsrc, Wsrc, h0, w0, Hin, Win,
                    (float)Hin / (float)Hout, (float)Win / (float)Wout,
                    (float)Hout / (float)Hin, (float)Wout / (float)Win, workspace};
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) {
        interp_bwd_rows_kernel<0><<<rs_grid(int64_t(N) * Hin * Wout), 256, 0, st>>>(a);
        interp_bwd_cols_kernel<0><<<rs_grid(int64_t(N) * Hsrc * Wsrc), 256, 0, st>>>(a);
    } else {
        interp_bwd_rows_kernel<1><<<rs_grid(int64_t(N) * Hin * Wout), 256, 0, st>>>(a);
        interp_bwd_cols_kernel<1><<<rs_grid(int64_t(N) * Hsrc * Wsrc), 256, 0, st>>>(a);
    }
    WM_LAUNCH_CHECK("wm_interp_bwd");
    return WM_OK;
}
