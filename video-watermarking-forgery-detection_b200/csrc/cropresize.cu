// Crop + resize back to full size (Crop.forward, noise_layers/crop.py:32-55: slice the crop
// rectangle, F.interpolate it to H x W, bilinear, align_corners=False) and its exact adjoint,
// for UP-scaling geometries (crop rates 0.5 .. 1 per axis -> source/destination slope in [0.45, 1]).
//
// The older tile kernels in resize.cu rebuilt their tap tables in every CTA and staged the source
// with per-element index arithmetic (issue-bound: 44 % / 16 % of the HBM roofline fwd / bwd).  Here:
//   * a small table kernel per call turns ATen's taps (same fp32 coordinate arithmetic, clamped
//     indices folded) into BANDED rows: start + NT weights per output (forward), start + BTT
//     weights per source sample (adjoint; BTT = 4 / 6 / 8 / 10 by slope, zero padded) — the crop box changes every call;
//   * the source region of a tile arrives by ONE TMA box load straight from the un-cropped frame
//     (box start = crop origin + band start, rounded down to the 16-byte boundary TMA needs for the
//     innermost coordinate — the taps are shifted by the remainder, so the rectangle itself is free);
//   * separable passes in shared memory: H pass lane = column (stride <= 1 word between lanes, no
//     bank conflicts), V pass lane = 4 (2) adjacent columns with LDS.128 (LDS.64) and broadcast
//     row weights, 128-bit (64-bit) coalesced stores;
//   * the adjoint walks tiles of the WHOLE source frame and writes the zeros outside the crop
//     rectangle itself (no memset pass), deterministic gather (ATen scatters with atomicAdd).
#include "interp_math.cuh"
#include "tma.cuh"
#include "wm_common.cuh"

namespace wm {
namespace {

constexpr int CR_THREADS = 256;
constexpr int CR_TW = 128, CR_TH = 32, CR_BW = 136, CR_BH = 36;     // forward: output tile, source box
constexpr int CA_TW = 64, CA_TH = 32, CA_BW = 160, CA_BH = 84;      // adjoint: source tile, cotangent box
constexpr float CR_SLOPE_MIN = 0.45f;

template <int MODE> struct CRK {
    static constexpr int NT = MODE == 0 ? 2 : 4;        // taps per output
};

inline int up4(int v) { return (v + 3) & ~3; }

struct CRTab { int lox, wx, loy, wy, flag, total; };
// forward tables: lo[Wout], w[Wout][NT], lo[Hout], w[Hout][NT]; adjoint: lo[Win], w[Win][BTT], lo[Hin], w[Hin][BTT]
inline CRTab cr_layout(int nx, int ny, int per) {
    CRTab t;
    t.lox = 0; t.wx = up4(nx); t.loy = t.wx + up4(nx * per); t.wy = t.loy + up4(ny);
    t.flag = t.wy + up4(ny * per); t.total = t.flag + 4;
    return t;
}

// ---- table kernels ------------------------------------------------------------------------------
template <int MODE>
__global__ void cr_fwd_tables_kernel(int* tab, CRTab L, int Hin, int Win, int Hout, int Wout, float sh, float sw) {
    constexpr int NT = CRK<MODE>::NT;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Wout + Hout) return;
    const bool isx = t < Wout;
    const int o = isx ? t : t - Wout, n_in = isx ? Win : Hin;
    int idx[4]; float w[4];
    taps<MODE>(isx ? sw : sh, o, n_in, idx, w);
    const int lo = min(idx[0], n_in - NT);
    float ww[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) ww[k] = 0.f;
#pragma unroll
    for (int k = 0; k < NT; ++k)
#pragma unroll
        for (int s = 0; s < NT; ++s) ww[s] += (idx[k] - lo == s) ? w[k] : 0.f;
    tab[(isx ? L.lox : L.loy) + o] = lo;
    float* wt = reinterpret_cast<float*>(tab) + (isx ? L.wx : L.wy) + o * NT;
#pragma unroll
    for (int k = 0; k < NT; ++k) wt[k] = ww[k];
}

template <int MODE>
__global__ void cr_adj_tables_kernel(int* tab, CRTab L, int Hin, int Win, int Hout, int Wout, float sh, float sw, int BTT) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Win + Hin) return;
    const bool isx = t < Win;
    const int i = isx ? t : t - Win, n_in = isx ? Win : Hin, n_out = isx ? Wout : Hout;
    const float s = isx ? sw : sh, inv = isx ? (float)Wout / (float)Win : (float)Hout / (float)Hin;
    int lo, hi;
    cand_range<MODE>(inv, i, n_in, n_out, lo, hi);
    while (lo <= hi && weight_of<MODE>(s, lo, n_in, i) == 0.f) ++lo;
    while (hi >= lo && weight_of<MODE>(s, hi, n_in, i) == 0.f) --hi;
    if (hi - lo + 1 > BTT) { tab[L.flag] = 1; hi = lo + BTT - 1; }       // cannot happen for slope >= 0.45
    const int start = max(min(lo, n_out - BTT), 0);
    tab[(isx ? L.lox : L.loy) + i] = start;
    float* wt = reinterpret_cast<float*>(tab) + (isx ? L.wx : L.wy) + i * BTT;
    for (int j = 0; j < BTT; ++j) {
        const int o = start + j;
        wt[j] = (o >= lo && o <= hi) ? weight_of<MODE>(s, o, n_in, i) : 0.f;
    }
}

// ---- forward --------------------------------------------------------------------------------------
struct CRArgs {
    const int* tab; CRTab L;
    float* y; int N, Hout, Wout, h0, w0;
    int bw, bh;                                    // TMA box of this geometry (<= CR_BW x CR_BH), = shared row pitch / rows
};

template <int MODE>
__global__ void __launch_bounds__(CR_THREADS) cr_fwd_kernel(const __grid_constant__ CUtensorMap tmap, const CRArgs a) {
    constexpr int NT = CRK<MODE>::NT;
    extern __shared__ __align__(128) float sm[];
    float* tile = sm;                              // [bh][bw] source region
    float* tmp = sm + a.bh * a.bw;                 // [bh][CR_TW] horizontally interpolated rows
    __shared__ uint64_t bar;
    __shared__ int s_lo[CR_TH];
    __shared__ float s_w[CR_TH][NT];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ox0 = blockIdx.x * CR_TW, oy0 = blockIdx.y * CR_TH, n = blockIdx.z;
    const float* wtab = reinterpret_cast<const float*>(a.tab);
    const int oxl = min(ox0 + CR_TW - 1, a.Wout - 1), oyl = min(oy0 + CR_TH - 1, a.Hout - 1);
    const int x_lo = __ldg(a.tab + a.L.lox + ox0), y_lo = __ldg(a.tab + a.L.loy + oy0);
    const int IH = __ldg(a.tab + a.L.loy + oyl) + NT - y_lo;
    // the box must start on a 16-byte boundary of the frame row: round the first column down, shift the taps
    const int xa = (a.w0 + x_lo) & ~3, xoff = a.w0 + x_lo - xa;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&bar, a.bw * a.bh * sizeof(float));
        tma_load_3d(tile, &tmap, xa, a.h0 + y_lo, n, &bar);
    }
    // band tables of this lane's four columns (lane, lane + 32, ...) and of the tile's rows
    int cx[4]; float cw[4][NT];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int ox = min(ox0 + lane + 32 * j, oxl);
        cx[j] = __ldg(a.tab + a.L.lox + ox) - x_lo + xoff;
#pragma unroll
        for (int k = 0; k < NT; ++k) cw[j][k] = __ldg(wtab + a.L.wx + ox * NT + k);
    }
    if (tid < CR_TH) {
        const int oy = min(oy0 + tid, oyl);
        s_lo[tid] = __ldg(a.tab + a.L.loy + oy) - y_lo;
#pragma unroll
        for (int k = 0; k < NT; ++k) s_w[tid][k] = __ldg(wtab + a.L.wy + oy * NT + k);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    // H pass over the staged source rows
    for (int r = warp; r < IH; r += CR_THREADS / 32) {
        const float* row = tile + r * a.bw;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float acc = cw[j][0] * row[cx[j]];
#pragma unroll
            for (int k = 1; k < NT; ++k) acc = fmaf(cw[j][k], row[cx[j] + k], acc);
            tmp[r * CR_TW + lane + 32 * j] = acc;
        }
    }
    __syncthreads();
    // V pass: four adjacent columns per lane
    const bool okc = ox0 + 4 * lane < a.Wout;
#pragma unroll
    for (int rr = warp; rr < CR_TH; rr += CR_THREADS / 32) {
        const int oy = oy0 + rr;
        if (oy >= a.Hout) break;
        const float* p = tmp + s_lo[rr] * CR_TW + 4 * lane;
        float4 v = *reinterpret_cast<const float4*>(p);
        const float w0 = s_w[rr][0];
        float4 acc = make_float4(w0 * v.x, w0 * v.y, w0 * v.z, w0 * v.w);
#pragma unroll
        for (int k = 1; k < NT; ++k) {
            v = *reinterpret_cast<const float4*>(p + k * CR_TW);
            const float wk = s_w[rr][k];
            acc.x = fmaf(wk, v.x, acc.x); acc.y = fmaf(wk, v.y, acc.y); acc.z = fmaf(wk, v.z, acc.z); acc.w = fmaf(wk, v.w, acc.w);
        }
        if (okc) stg128(a.y + (int64_t(n) * a.Hout + oy) * a.Wout + ox0 + 4 * lane, acc);
    }
}

// ---- adjoint --------------------------------------------------------------------------------------
struct CAArgs {
    const int* tab; CRTab L;
    float* gx; int N, Hsrc, Wsrc, h0, w0, Hin, Win;
    int bw, bh;                                    // TMA box of this geometry (<= CA_BW x CA_BH)
};

template <int BTT>
__global__ void __launch_bounds__(CR_THREADS) cr_adj_kernel(const __grid_constant__ CUtensorMap tmap, const CAArgs a) {
    extern __shared__ __align__(128) float sm[];
    float* G = sm;                                 // [bh][bw] cotangent region
    float* tmp = sm + a.bh * a.bw;                 // [bh][CA_TW]
    __shared__ uint64_t bar;
    __shared__ int s_lo[CA_TH];
    __shared__ float s_w[CA_TH][BTT];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sx0 = blockIdx.x * CA_TW, sy0 = blockIdx.y * CA_TH, n = blockIdx.z;
    float* dst = a.gx + (int64_t(n) * a.Hsrc + sy0) * a.Wsrc + sx0 + 2 * lane;
    const bool okc = sx0 + 2 * lane < a.Wsrc;
    // the part of this tile that lies inside the crop rectangle, in rectangle coordinates
    const int ix_a = max(sx0 - a.w0, 0), ix_b = min(sx0 + CA_TW - 1 - a.w0, a.Win - 1);
    const int iy_a = max(sy0 - a.h0, 0), iy_b = min(sy0 + CA_TH - 1 - a.h0, a.Hin - 1);
    if (ix_a > ix_b || iy_a > iy_b) {              // wholly outside: the gradient is zero there
        for (int rr = warp; rr < CA_TH && sy0 + rr < a.Hsrc; rr += CR_THREADS / 32)
            if (okc) *reinterpret_cast<float2*>(dst + int64_t(rr) * a.Wsrc) = make_float2(0.f, 0.f);
        return;
    }
    const float* wtab = reinterpret_cast<const float*>(a.tab);
    const int g_xlo = __ldg(a.tab + a.L.lox + ix_a), g_ylo = __ldg(a.tab + a.L.loy + iy_a);
    const int GH = __ldg(a.tab + a.L.loy + iy_b) + BTT - g_ylo;
    const int gxa = g_xlo & ~3, goff = g_xlo - gxa;       // 16-byte aligned box start, taps shifted
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&bar, a.bw * a.bh * sizeof(float));
        tma_load_3d(G, &tmap, gxa, g_ylo, n, &bar);
    }
    int cx[2]; float cw[2][BTT];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int ix = sx0 + lane + 32 * j - a.w0;
        const bool in = ix >= 0 && ix < a.Win;
        const int ic = min(max(ix, ix_a), ix_b);
        cx[j] = __ldg(a.tab + a.L.lox + ic) - g_xlo + goff;
#pragma unroll
        for (int k = 0; k < BTT; ++k) cw[j][k] = in ? __ldg(wtab + a.L.wx + ic * BTT + k) : 0.f;
    }
    if (tid < CA_TH) {
        const int iy = sy0 + tid - a.h0;
        const bool in = iy >= 0 && iy < a.Hin;
        const int ic = min(max(iy, iy_a), iy_b);
        s_lo[tid] = __ldg(a.tab + a.L.loy + ic) - g_ylo;
#pragma unroll
        for (int k = 0; k < BTT; ++k) s_w[tid][k] = in ? __ldg(wtab + a.L.wy + ic * BTT + k) : 0.f;
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int r = warp; r < GH; r += CR_THREADS / 32) {
        const float* row = G + r * a.bw;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float acc = cw[j][0] * row[cx[j]];
#pragma unroll
            for (int k = 1; k < BTT; ++k) acc = fmaf(cw[j][k], row[cx[j] + k], acc);
            tmp[r * CA_TW + lane + 32 * j] = acc;
        }
    }
    __syncthreads();
#pragma unroll
    for (int rr = warp; rr < CA_TH; rr += CR_THREADS / 32) {
        if (sy0 + rr >= a.Hsrc) break;
        const float* p = tmp + s_lo[rr] * CA_TW + 2 * lane;
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < BTT; ++k) {
            const float2 v = *reinterpret_cast<const float2*>(p + k * CA_TW);
            const float wk = s_w[rr][k];
            acc.x = fmaf(wk, v.x, acc.x); acc.y = fmaf(wk, v.y, acc.y);
        }
        if (okc) *reinterpret_cast<float2*>(dst + int64_t(rr) * a.Wsrc) = acc;
    }
}

bool cr_geometry_ok(int Hin, int Win, int Hout, int Wout, int N, int mode) {
    if (mode != 0 && mode != 1) return false;
    const int nt = mode == 0 ? 2 : 4, btt = mode == 0 ? 6 : 10;
    if (N <= 0 || N > 65535 || Hin < nt || Win < nt || Hout < btt || Wout < btt || Wout % 4) return false;
    const float sh = (float)Hin / (float)Hout, sw = (float)Win / (float)Wout;
    return sh >= CR_SLOPE_MIN && sh <= 1.f && sw >= CR_SLOPE_MIN && sw <= 1.f && (Hout + CR_TH - 1) / CR_TH <= 65535;
}

}  // namespace
}  // namespace wm

using namespace wm;

// 1 if wm_cropresize_fwd / _bwd serve this geometry (up-scaling crop, slopes in [0.45, 1], Wout % 4 == 0)
extern "C" int wm_cropresize_ok(int Hin, int Win, int Hout, int Wout, int N, int mode) {
    return cr_geometry_ok(Hin, Win, Hout, Wout, N, mode) && tmap_encoder() != nullptr ? 1 : 0;
}

// 4-byte words of table workspace either direction needs
extern "C" int64_t wm_cropresize_table_words(int Hin, int Win, int Hout, int Wout, int mode) {
    const int nt = mode == 0 ? 2 : 4, btt = mode == 0 ? 6 : 10;
    const int f = cr_layout(Wout, Hout, nt).total, b = cr_layout(Win, Hin, btt).total;
    return f > b ? f : b;
}

extern "C" int wm_cropresize_fwd(const float* x, int64_t x_sp, int64_t x_sh, int Hsrc, int Wsrc, int h0, int w0, int Hin, int Win,
                                 float* y, int N, int Hout, int Wout, int mode, int32_t* tables, void* stream) {
    if (N == 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(x && y && tables, WM_E_NULL, "wm_cropresize_fwd: null pointer");
    WM_REQUIRE(cr_geometry_ok(Hin, Win, Hout, Wout, N, mode), WM_E_SHAPE,
               "wm_cropresize_fwd: geometry %dx%d -> %dx%d (mode %d) is outside the fast path, use wm_interp_fwd", Hin, Win, Hout, Wout, mode);
    WM_REQUIRE(h0 >= 0 && w0 >= 0 && h0 + Hin <= Hsrc && w0 + Win <= Wsrc, WM_E_SHAPE, "wm_cropresize_fwd: rectangle outside the frame");
    WM_REQUIRE(aligned(y, 16) && aligned(tables, 16) && tmap_ok(x, x_sp, x_sh, 4), WM_E_ALIGN,
               "wm_cropresize_fwd: x / y / tables must be 16-byte aligned with 16-byte row and plane strides");
    cudaStream_t st = (cudaStream_t)stream;
    const float sh = (float)Hin / (float)Hout, sw = (float)Win / (float)Wout;
    const CRTab L = cr_layout(Wout, Hout, mode == 0 ? 2 : 4);
    // box sized for THIS slope: band starts advance by at most slope * 127 + 1 over a tile, + taps + alignment shift
    const int nt = mode == 0 ? 2 : 4;
    int bw = up4((int)ceilf((CR_TW - 1) * sw) + 1 + nt + 3 + 1), bh = (int)ceilf((CR_TH - 1) * sh) + 1 + nt + 1;
    bw = bw > CR_BW ? CR_BW : bw; bh = bh > CR_BH ? CR_BH : bh;
    CUtensorMap tm;
    if (int rc = tmap_planes(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, N, Hsrc, Wsrc, x_sp, x_sh, bw, bh)) {
        set_error("wm_cropresize_fwd: cuTensorMapEncodeTiled failed (%d)", rc);
        return WM_E_ARG;
    }
    CRArgs a{tables, L, y, N, Hout, Wout, h0, w0, bw, bh};
    const size_t smem = sizeof(float) * (size_t(bh) * bw + size_t(bh) * CR_TW);
    const dim3 grid((Wout + CR_TW - 1) / CR_TW, (Hout + CR_TH - 1) / CR_TH, N);
    const int tb = (Wout + Hout + 127) / 128;
    if (mode == 0) {
        cr_fwd_tables_kernel<0><<<tb, 128, 0, st>>>(tables, L, Hin, Win, Hout, Wout, sh, sw);
        cr_fwd_kernel<0><<<grid, CR_THREADS, smem, st>>>(tm, a);
    } else {
        cr_fwd_tables_kernel<1><<<tb, 128, 0, st>>>(tables, L, Hin, Win, Hout, Wout, sh, sw);
        cr_fwd_kernel<1><<<grid, CR_THREADS, smem, st>>>(tm, a);
    }
    WM_LAUNCH_CHECK("wm_cropresize_fwd");
    return WM_OK;
}

extern "C" int wm_cropresize_bwd(const float* gy, float* gx, int Hsrc, int Wsrc, int h0, int w0, int Hin, int Win,
                                 int N, int Hout, int Wout, int mode, int32_t* tables, void* stream) {
    if (N == 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(gy && gx && tables, WM_E_NULL, "wm_cropresize_bwd: null pointer");
    WM_REQUIRE(cr_geometry_ok(Hin, Win, Hout, Wout, N, mode) && Wsrc % 2 == 0, WM_E_SHAPE,
               "wm_cropresize_bwd: geometry %dx%d -> %dx%d (mode %d) is outside the fast path, use wm_interp_bwd", Hin, Win, Hout, Wout, mode);
    WM_REQUIRE(h0 >= 0 && w0 >= 0 && h0 + Hin <= Hsrc && w0 + Win <= Wsrc && (Hsrc + CA_TH - 1) / CA_TH <= 65535, WM_E_SHAPE,
               "wm_cropresize_bwd: rectangle outside the frame");
    WM_REQUIRE(aligned(gy, 16) && aligned(gx, 8) && aligned(tables, 16), WM_E_ALIGN, "wm_cropresize_bwd: gy / tables 16-byte, gx 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const float sh = (float)Hin / (float)Hout, sw = (float)Win / (float)Wout;
    // outputs touching one source sample: at most floor(taps / slope) + 1 -> pad the adjoint bands to 4 / 6 / 8 / 10
    const float smin = sh < sw ? sh : sw;
    const int need = (int)floorf((mode == 0 ? 2.f : 4.f) / smin) + 1;
    const int btt = need <= 4 ? 4 : (need <= 6 ? 6 : (need <= 8 ? 8 : 10));
    int bw = up4((int)ceilf((CA_TW - 1) / sw) + 1 + btt + 3 + 1), bh = (int)ceilf((CA_TH - 1) / sh) + 1 + btt + 1;
    bw = bw > CA_BW ? CA_BW : bw; bh = bh > CA_BH ? CA_BH : bh;
    const CRTab L = cr_layout(Win, Hin, btt);
    CUtensorMap tm;
    if (int rc = tmap_planes(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, gy, N, Hout, Wout, int64_t(Hout) * Wout, Wout, bw, bh)) {
        set_error("wm_cropresize_bwd: cuTensorMapEncodeTiled failed (%d)", rc);
        return WM_E_ARG;
    }
    CAArgs a{tables, L, gx, N, Hsrc, Wsrc, h0, w0, Hin, Win, bw, bh};
    const size_t smem = sizeof(float) * (size_t(bh) * bw + size_t(bh) * CA_TW);
    const dim3 grid((Wsrc + CA_TW - 1) / CA_TW, (Hsrc + CA_TH - 1) / CA_TH, N);
    const int tb = (Win + Hin + 127) / 128;
    cudaError_t e = cudaMemsetAsync(tables + L.flag, 0, 4 * sizeof(int32_t), st);
    if (e != cudaSuccess) return cuda_fail(e, "wm_cropresize_bwd");
    if (mode == 0) cr_adj_tables_kernel<0><<<tb, 128, 0, st>>>(tables, L, Hin, Win, Hout, Wout, sh, sw, btt);
    else cr_adj_tables_kernel<1><<<tb, 128, 0, st>>>(tables, L, Hin, Win, Hout, Wout, sh, sw, btt);
    auto kern = btt == 4 ? cr_adj_kernel<4> : (btt == 6 ? cr_adj_kernel<6> : (btt == 8 ? cr_adj_kernel<8> : cr_adj_kernel<10>));
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "wm_cropresize_bwd");
    kern<<<grid, CR_THREADS, smem, st>>>(tm, a);
    WM_LAUNCH_CHECK("wm_cropresize_bwd");
    return WM_OK;
}
