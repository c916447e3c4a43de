// Bilinear / bicubic interpolation with F.interpolate(size=..., align_corners=False) semantics
// (ATen upsample_bilinear2d / upsample_bicubic2d, A = -0.75), forward and exact transpose.
//
// Replaces: Resize.forward (noise_layers/resize.py:38-53: two F.interpolate + clamp) and
// Crop.forward (noise_layers/crop.py:48-53: slice + bilinear F.interpolate).  The source window
// arguments let Crop read the crop rectangle in place; `clamp01` fuses Resize's clamp and can
// emit the clamp mask as a bit plane (1 bit/value) so that the backward needs no recompute.
//
// Both directions are SEPARABLE TILE kernels: a CTA owns a 32x64 tile of its output, builds the
// per-column and per-row tap tables (index + weights, computed with the same fp32 coordinate
// arithmetic as ATen) once in shared memory, stages the source region it needs, runs the
// horizontal pass into a second shared buffer and the vertical pass to global memory.
// The backward is the exact adjoint as a GATHER (each input pixel sums the outputs whose taps
// touch it), so it is deterministic — ATen's upsample backward uses atomicAdd.
// Scales outside [0.4, 2.2] (never produced by Resize's (0.5, 1.5) range) use the simple
// per-element kernels at the bottom.
#include "wm_common.cuh"
#include "interp_math.cuh"

namespace wm {

constexpr int RS_TH = 32, RS_TW = 64, RS_THREADS = 256;
constexpr int RS_MAXC = 16;                  // max outputs touching one input sample (adjoint tables)
constexpr float RS_SCALE_MIN = 0.4f, RS_SCALE_MAX = 2.2f;

struct InterpArgs {
    const float* x; int64_t x_sp, x_sh; int h0, w0, Hin, Win;
    float* y; int N, Hout, Wout; float sh, sw; int clamp01;
    uint32_t* maskbits; int mask_words_per_row;
    int IHmax, IWmax;                      // shared-memory region bounds (tiled kernels)
};

// =============================================================================================
// forward, tiled
// =============================================================================================
template <int MODE>
__global__ void __launch_bounds__(RS_THREADS) interp_fwd_tiled_kernel(const InterpArgs a) {
    constexpr int NT = MODE == 0 ? 2 : 4;
    extern __shared__ float sm[];
    __shared__ int   cIdx[RS_TW][4];  __shared__ float cW[RS_TW][4];
    __shared__ int   rIdx[RS_TH][4];  __shared__ float rW[RS_TH][4];
    const int IWp = a.IWmax | 1;                  // odd row pitch
    float* tile = sm;                             // [IHmax][IWp]
    float* tmp = sm + a.IHmax * IWp;              // [IHmax][RS_TW]
    const int ox0 = blockIdx.x * RS_TW, oy0 = blockIdx.y * RS_TH, n = blockIdx.z;
    const int t = threadIdx.x;
    if (t < RS_TW) {
        const int ox = min(ox0 + t, a.Wout - 1);
        int idx[4]; float w[4];
        taps<MODE>(a.sw, ox, a.Win, idx, w);
#pragma unroll
        for (int k = 0; k < 4; ++k) { cIdx[t][k] = idx[k]; cW[t][k] = w[k]; }
    } else if (t < RS_TW + RS_TH) {
        const int r = t - RS_TW;
        const int oy = min(oy0 + r, a.Hout - 1);
        int idx[4]; float w[4];
        taps<MODE>(a.sh, oy, a.Hin, idx, w);
#pragma unroll
        for (int k = 0; k < 4; ++k) { rIdx[r][k] = idx[k]; rW[r][k] = w[k]; }
    }
    __syncthreads();
    const int cx_lo = cIdx[0][0], cx_hi = cIdx[RS_TW - 1][NT - 1];
    const int ry_lo = rIdx[0][0], ry_hi = rIdx[RS_TH - 1][NT - 1];
    const int IW = cx_hi - cx_lo + 1, IH = ry_hi - ry_lo + 1;
    const float* src = a.x + int64_t(n) * a.x_sp + int64_t(a.h0 + ry_lo) * a.x_sh + (a.w0 + cx_lo);
    for (int i = t; i < IH * IW; i += RS_THREADS) {
        const int r = i / IW, c = i - r * IW;
        tile[r * IWp + c] = __ldg(src + int64_t(r) * a.x_sh + c);
    }
    __syncthreads();
    {   // horizontal pass: thread owns output column c, walks the staged rows
        const int c = t % RS_TW;
        int id[4]; float w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { id[k] = cIdx[c][k] - cx_lo; w[k] = cW[c][k]; }
        for (int r = t / RS_TW; r < IH; r += RS_THREADS / RS_TW) {
            const float* row = tile + r * IWp;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < NT; ++k) acc = fmaf(w[k], row[id[k]], acc);
            tmp[r * RS_TW + c] = acc;
        }
    }
    __syncthreads();
    {   // vertical pass
        const int c = t % RS_TW, ox = ox0 + c;
        float* dst = a.y + int64_t(n) * a.Hout * a.Wout;
        for (int r = t / RS_TW; r < RS_TH; r += RS_THREADS / RS_TW) {
            const int oy = oy0 + r;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < NT; ++k) acc = fmaf(rW[r][k], tmp[(rIdx[r][k] - ry_lo) * RS_TW + c], acc);
            const bool ok = oy < a.Hout && ox < a.Wout;
            bool inside = true;
            if (a.clamp01) { inside = acc >= 0.f && acc <= 1.f; acc = clamp01_nan(acc); }
            if (ok) dst[int64_t(oy) * a.Wout + ox] = acc;
            if (a.maskbits) {
                const unsigned bits = __ballot_sync(0xffffffffu, inside && ok);
                if ((t & 31) == 0 && oy < a.Hout && ox < a.Wout)
                    a.maskbits[(int64_t(n) * a.Hout + oy) * a.mask_words_per_row + (ox >> 5)] = bits;
            }
        }
    }
}

// =============================================================================================
// adjoint, tiled:  gx[window] = W_h^T (gy .* mask) W_w
// =============================================================================================
struct InterpBwdArgs {
    const float* gy; const float* pre; const uint32_t* maskbits; int mask_words_per_row;
    int N, Hout, Wout;
    float* gx; int Hsrc, Wsrc, h0, w0, Hin, Win;
    float sh, sw, ish, isw; float* ws;
    int GHmax, GWmax;
};

template <int MODE>
__global__ void __launch_bounds__(RS_THREADS) interp_adj_tiled_kernel(const InterpBwdArgs a) {
    extern __shared__ float sm[];
    __shared__ int   cLo[RS_TW], cCnt[RS_TW];  __shared__ float cW[RS_TW][RS_MAXC];
    __shared__ int   rLo[RS_TH], rCnt[RS_TH];  __shared__ float rW[RS_TH][RS_MAXC];
    __shared__ int   red[4];
    const int GWp = a.GWmax | 1;
    float* G = sm;                                // [GHmax][GWp]   staged (masked) cotangent region
    float* tmp = sm + a.GHmax * GWp;              // [GHmax][RS_TW]
    const int ix0 = blockIdx.x * RS_TW, iy0 = blockIdx.y * RS_TH, n = blockIdx.z;
    const int t = threadIdx.x;
    if (t < RS_TW) {
        const int ix = ix0 + t;
        int lo = 0, cnt = 0;
        if (ix < a.Win) {
            int hi;
            cand_range<MODE>(a.isw, ix, a.Win, a.Wout, lo, hi);
            // trim to the outputs that really touch ix, keep at most RS_MAXC
            while (lo <= hi && weight_of<MODE>(a.sw, lo, a.Win, ix) == 0.f) ++lo;
            while (hi >= lo && weight_of<MODE>(a.sw, hi, a.Win, ix) == 0.f) --hi;
            cnt = min(hi - lo + 1, RS_MAXC);
            for (int j = 0; j < cnt; ++j) cW[t][j] = weight_of<MODE>(a.sw, lo + j, a.Win, ix);
        }
        cLo[t] = lo; cCnt[t] = max(cnt, 0);
    } else if (t < RS_TW + RS_TH) {
        const int r = t - RS_TW, iy = iy0 + r;
        int lo = 0, cnt = 0;
        if (iy < a.Hin) {
            int hi;
            cand_range<MODE>(a.ish, iy, a.Hin, a.Hout, lo, hi);
            while (lo <= hi && weight_of<MODE>(a.sh, lo, a.Hin, iy) == 0.f) ++lo;
            while (hi >= lo && weight_of<MODE>(a.sh, hi, a.Hin, iy) == 0.f) --hi;
            cnt = min(hi - lo + 1, RS_MAXC);
            for (int j = 0; j < cnt; ++j) rW[r][j] = weight_of<MODE>(a.sh, lo + j, a.Hin, iy);
        }
        rLo[r] = lo; rCnt[r] = max(cnt, 0);
    }
    __syncthreads();
    if (t == 0) {
        int glo = 1 << 30, ghi = -1, hlo = 1 << 30, hhi = -1;
        for (int c = 0; c < RS_TW; ++c) if (cCnt[c] > 0) { glo = min(glo, cLo[c]); ghi = max(ghi, cLo[c] + cCnt[c] - 1); }
        for (int r = 0; r < RS_TH; ++r) if (rCnt[r] > 0) { hlo = min(hlo, rLo[r]); hhi = max(hhi, rLo[r] + rCnt[r] - 1); }
        red[0] = glo; red[1] = ghi; red[2] = hlo; red[3] = hhi;
    }
    __syncthreads();
    const int gx_lo = red[0], gx_hi = red[1], gy_lo = red[2], gy_hi = red[3];
    const int GW = max(gx_hi - gx_lo + 1, 0), GH = max(gy_hi - gy_lo + 1, 0);
    const float* gsrc = a.gy + int64_t(n) * a.Hout * a.Wout;
    for (int i = t; i < GH * GW; i += RS_THREADS) {
        const int r = i / GW, c = i - r * GW;
        const int oy = gy_lo + r, ox = gx_lo + c;
        float g = __ldg(gsrc + int64_t(oy) * a.Wout + ox);
        if (a.maskbits) {
            const uint32_t wbits = __ldg(a.maskbits + (int64_t(n) * a.Hout + oy) * a.mask_words_per_row + (ox >> 5));
            g = ((wbits >> (ox & 31)) & 1u) ? g : 0.f;
        } else if (a.pre) {
            const float p = __ldg(a.pre + int64_t(n) * a.Hout * a.Wout + int64_t(oy) * a.Wout + ox);
            g = (p >= 0.f && p <= 1.f) ? g : 0.f;
        }
        G[r * GWp + c] = g;
    }
    __syncthreads();
    {   // horizontal adjoint
        const int c = t % RS_TW;
        const int lo = cLo[c] - gx_lo, cnt = cCnt[c];
        for (int r = t / RS_TW; r < GH; r += RS_THREADS / RS_TW) {
            const float* row = G + r * GWp + lo;
            float acc = 0.f;
            for (int j = 0; j < cnt; ++j) acc = fmaf(cW[c][j], row[j], acc);
            tmp[r * RS_TW + c] = acc;
        }
    }
    __syncthreads();
    {   // vertical adjoint
        const int c = t % RS_TW, ix = ix0 + c;
        float* dst = a.gx + int64_t(n) * a.Hsrc * a.Wsrc + int64_t(a.h0) * a.Wsrc + a.w0;
        for (int r = t / RS_TW; r < RS_TH; r += RS_THREADS / RS_TW) {
            const int iy = iy0 + r;
            if (iy >= a.Hin || ix >= a.Win) continue;
            const int lo = rLo[r] - gy_lo, cnt = rCnt[r];
            float acc = 0.f;
            for (int j = 0; j < cnt; ++j) acc = fmaf(rW[r][j], tmp[(lo + j) * RS_TW + c], acc);
            dst[int64_t(iy) * a.Wsrc + ix] = acc;
        }
    }
}

// =============================================================================================
// simple per-element kernels (any scale)
// =============================================================================================
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
        float tmpv[4];
#pragma unroll
        for (int r = 0; r < NT; ++r) {
            const float* row = src + int64_t(iy[r]) * a.x_sh;
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < NT; ++c) s = fmaf(wx[c], __ldg(row + ix[c]), s);
            tmpv[r] = s;
        }
#pragma unroll
        for (int r = 0; r < NT; ++r) acc = fmaf(wy[r], tmpv[r], acc);
        bool inside = true;
        if (a.clamp01) { inside = acc >= 0.f && acc <= 1.f; acc = clamp01_nan(acc); }
        a.y[i] = acc;
        if (a.maskbits && !inside)
            atomicAnd(a.maskbits + (int64_t(n) * a.Hout + oy) * a.mask_words_per_row + (ox >> 5), ~(1u << (ox & 31)));
    }
}

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
                if (a.maskbits) {
                    const uint32_t wb = __ldg(a.maskbits + (int64_t(n) * a.Hout + oy) * a.mask_words_per_row + (ox >> 5));
                    g = ((wb >> (ox & 31)) & 1u) ? g : 0.f;
                } else if (a.pre) {
                    const float p = __ldg(a.pre + base + int64_t(oy) * a.Wout);
                    g = (p >= 0.f && p <= 1.f) ? g : 0.f;
                }
                acc = fmaf(w, g, acc);
            }
        }
        a.ws[i] = acc;
    }
}

// pass 2: gx[n][h0+iy][w0+ix] = sum_ox Wx(ox, ix) * ws[n][iy][ox]  (window only)
template <int MODE>
__global__ void __launch_bounds__(256) interp_bwd_cols_kernel(const InterpBwdArgs a) {
    const int64_t total = int64_t(a.N) * a.Hin * a.Win;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int ix = int(i % a.Win), iy = int((i / a.Win) % a.Hin), n = int(i / (int64_t(a.Win) * a.Hin));
        int lo, hi;
        cand_range<MODE>(a.isw, ix, a.Win, a.Wout, lo, hi);
        const float* row = a.ws + (int64_t(n) * a.Hin + iy) * a.Wout;
        float acc = 0.f;
        for (int ox = lo; ox <= hi; ++ox) {
            const float w = weight_of<MODE>(a.sw, ox, a.Win, ix);
            if (w != 0.f) acc = fmaf(w, row[ox], acc);
        }
        a.gx[int64_t(n) * a.Hsrc * a.Wsrc + int64_t(a.h0 + iy) * a.Wsrc + a.w0 + ix] = acc;
    }
}

static inline unsigned rs_grid(int64_t total) {
    const int64_t want = (total + 255) / 256, cap = int64_t(sm_count()) * 32;
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

static inline bool rs_tiled_ok(float sh, float sw, int N) {
    return sh >= RS_SCALE_MIN && sh <= RS_SCALE_MAX && sw >= RS_SCALE_MIN && sw <= RS_SCALE_MAX && N <= 65535;
}

}  // namespace wm

using namespace wm;

extern "C" int wm_interp_fwd(const float* x, int64_t x_sp, int64_t x_sh, int Hsrc, int Wsrc, int h0, int w0, int Hin, int Win,
                             float* y, int N, int Hout, int Wout, int mode, int clamp01,
                             uint32_t* maskbits, void* stream) {
    if (N == 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(x && y, WM_E_NULL, "wm_interp_fwd: null pointer");
    WM_REQUIRE(mode == 0 || mode == 1, WM_E_ARG, "wm_interp_fwd: mode must be 0 (bilinear) or 1 (bicubic)");
    WM_REQUIRE(N >= 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0 && h0 >= 0 && w0 >= 0, WM_E_SHAPE,
               "wm_interp_fwd: bad shape N=%d in=%dx%d out=%dx%d", N, Hin, Win, Hout, Wout);
    WM_REQUIRE(int64_t(h0) + Hin <= Hsrc && int64_t(w0) + Win <= Wsrc, WM_E_SHAPE,
               "wm_interp_fwd: window [%d:%d, %d:%d] is outside the %dx%d source plane (it is read in place)",
               h0, h0 + Hin, w0, w0 + Win, Hsrc, Wsrc);
    if (N == 0) return WM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    InterpArgs a{};
    a.x = x; a.x_sp = x_sp; a.x_sh = x_sh; a.h0 = h0; a.w0 = w0; a.Hin = Hin; a.Win = Win;
    a.y = y; a.N = N; a.Hout = Hout; a.Wout = Wout;
    a.sh = (float)Hin / (float)Hout; a.sw = (float)Win / (float)Wout; a.clamp01 = clamp01;
    a.maskbits = maskbits; a.mask_words_per_row = (Wout + 31) / 32;
    if (rs_tiled_ok(a.sh, a.sw, N)) {
        a.IHmax = (int)(RS_TH * a.sh) + 6; a.IWmax = (int)(RS_TW * a.sw) + 6;
        const size_t smem = sizeof(float) * (size_t(a.IHmax) * (a.IWmax | 1) + size_t(a.IHmax) * RS_TW);
        dim3 grid((Wout + RS_TW - 1) / RS_TW, (Hout + RS_TH - 1) / RS_TH, N);
        WM_REQUIRE(grid.y <= 65535, WM_E_SHAPE, "wm_interp_fwd: output too tall");
        auto kern = mode == 0 ? interp_fwd_tiled_kernel<0> : interp_fwd_tiled_kernel<1>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "wm_interp_fwd");
        kern<<<grid, RS_THREADS, smem, st>>>(a);
    } else {
        if (maskbits) {
            cudaError_t e = cudaMemsetAsync(maskbits, 0xff, sizeof(uint32_t) * size_t(N) * Hout * a.mask_words_per_row, st);
            if (e != cudaSuccess) return cuda_fail(e, "wm_interp_fwd(mask)");
        }
        const unsigned grid = rs_grid(int64_t(N) * Hout * Wout);
        if (mode == 0) interp_fwd_kernel<0><<<grid, 256, 0, st>>>(a);
        else interp_fwd_kernel<1><<<grid, 256, 0, st>>>(a);
    }
    WM_LAUNCH_CHECK("wm_interp_fwd");
    return WM_OK;
}

extern "C" int wm_interp_bwd(const float* gy, const float* pre, const uint32_t* maskbits, int N, int Hout, int Wout,
                             float* gx, int Hsrc, int Wsrc, int h0, int w0, int Hin, int Win,
                             int mode, float* workspace, void* stream) {
    if (N == 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(gy && gx, WM_E_NULL, "wm_interp_bwd: null pointer");
    WM_REQUIRE(mode == 0 || mode == 1, WM_E_ARG, "wm_interp_bwd: mode must be 0 (bilinear) or 1 (bicubic)");
    WM_REQUIRE(N >= 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0 && h0 >= 0 && w0 >= 0 &&
               h0 + Hin <= Hsrc && w0 + Win <= Wsrc, WM_E_SHAPE,
               "wm_interp_bwd: bad shape N=%d window=%dx%d@(%d,%d) src=%dx%d out=%dx%d", N, Hin, Win, h0, w0, Hsrc, Wsrc, Hout, Wout);
    if (N == 0) return WM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    InterpBwdArgs a{};
    a.gy = gy; a.pre = pre; a.maskbits = maskbits; a.mask_words_per_row = (Wout + 31) / 32;
    a.N = N; a.Hout = Hout; a.Wout = Wout; a.gx = gx; a.Hsrc = Hsrc; a.Wsrc = Wsrc;
    a.h0 = h0; a.w0 = w0; a.Hin = Hin; a.Win = Win;
    a.sh = (float)Hin / (float)Hout; a.sw = (float)Win / (float)Wout;
    a.ish = (float)Hout / (float)Hin; a.isw = (float)Wout / (float)Win; a.ws = workspace;
    if (Hin != Hsrc || Win != Wsrc) {      // gradient is zero outside the source window
        cudaError_t e = cudaMemsetAsync(gx, 0, sizeof(float) * size_t(N) * Hsrc * Wsrc, st);
        if (e != cudaSuccess) return cuda_fail(e, "wm_interp_bwd(memset)");
    }
    if (rs_tiled_ok(a.sh, a.sw, N)) {
        const int reach = mode == 0 ? 1 : 2;
        a.GHmax = (int)((RS_TH + 2 * reach + 1) * a.ish) + 8; a.GWmax = (int)((RS_TW + 2 * reach + 1) * a.isw) + 8;
        const size_t smem = sizeof(float) * (size_t(a.GHmax) * (a.GWmax | 1) + size_t(a.GHmax) * RS_TW);
        dim3 grid((Win + RS_TW - 1) / RS_TW, (Hin + RS_TH - 1) / RS_TH, N);
        WM_REQUIRE(grid.y <= 65535, WM_E_SHAPE, "wm_interp_bwd: input too tall");
        auto kern = mode == 0 ? interp_adj_tiled_kernel<0> : interp_adj_tiled_kernel<1>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "wm_interp_bwd");
        kern<<<grid, RS_THREADS, smem, st>>>(a);
    } else {
        WM_REQUIRE(workspace != nullptr, WM_E_NULL,
                   "wm_interp_bwd: scale outside [0.4, 2.2] needs a workspace of N*Hin*Wout floats");
        if (mode == 0) {
            interp_bwd_rows_kernel<0><<<rs_grid(int64_t(N) * Hin * Wout), 256, 0, st>>>(a);
            interp_bwd_cols_kernel<0><<<rs_grid(int64_t(N) * Hin * Win), 256, 0, st>>>(a);
        } else {
            interp_bwd_rows_kernel<1><<<rs_grid(int64_t(N) * Hin * Wout), 256, 0, st>>>(a);
            interp_bwd_cols_kernel<1><<<rs_grid(int64_t(N) * Hin * Win), 256, 0, st>>>(a);
        }
    }
    WM_LAUNCH_CHECK("wm_interp_bwd");
    return WM_OK;
}

// 1 if the tiled kernels serve this geometry (then wm_interp_bwd needs no workspace)
extern "C" int wm_interp_is_tiled(int Hin, int Win, int Hout, int Wout, int N) {
    return rs_tiled_ok((float)Hin / (float)Hout, (float)Win / (float)Wout, N) ? 1 : 0;
}
