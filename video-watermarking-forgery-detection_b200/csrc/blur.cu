// Separable Gaussian blur over N = B*C planes (depth-wise), zero or reflect border.
//
// Replaces: GaussianBlur.forward (noise_layers/gaussian_blur.py:53-56), which builds a fresh
// nn.Conv2d and uploads its weights on EVERY call, and GF / kornia GaussianBlur2d
// (noise_layers/gaussian_filter.py:9-13).  The 2-D kernel of the reference is the outer product
// of normalised 1-D taps, so one CTA stages a (TH+2r) x (TW+2r) halo tile in shared memory,
// runs the horizontal pass into a second shared buffer and the vertical pass straight to
// global memory: x is read once (+halo) and y written once.
// Backward: the filter is symmetric, so with zero padding the adjoint is the same operator.
// For the reflect border the adjoint folds the out-of-range part of the zero-extended
// response back onto the interior (gather form, deterministic).
#include "tma.cuh"
#include "wm_common.cuh"

namespace wm {

constexpr int BL_TW = 128, BL_TH = 32, BL_THREADS = 256, BL_MAXK = 31;

struct BlurArgs {
    const float* x; int64_t x_sp, x_sh;
    float* y; int N, H, W, k, r, border;
    float taps[BL_MAXK + 1];
};

__device__ __forceinline__ int reflect_idx(int i, int n) {
    // kornia / F.pad(mode='reflect'): -i -> i, n-1+i -> n-1-i   (requires r < n)
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

__global__ void __launch_bounds__(BL_THREADS) gaussblur_kernel(const BlurArgs a) {
    extern __shared__ float sm[];
    const int r = a.r, k = a.k;
    const int IW = BL_TW + 2 * r, IH = BL_TH + 2 * r;
    float* tin = sm;                       // [IH][IW]
    float* tmp = sm + IH * IW;             // [IH][BL_TW]
    const int tiles_x = (a.W + BL_TW - 1) / BL_TW;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int n = blockIdx.y;
    const int x0 = tx * BL_TW, y0 = ty * BL_TH;
    const float* src = a.x + int64_t(n) * a.x_sp;
    for (int i = threadIdx.x; i < IH * IW; i += BL_THREADS) {
        const int ly = i / IW, lx = i - ly * IW;
        int gy = y0 + ly - r, gx = x0 + lx - r;
        float v = 0.f;
        if (a.border == 1) {
            // rows/cols beyond the reflected range of this (possibly partial) tile are unused
            gy = reflect_idx(gy, a.H); gx = reflect_idx(gx, a.W);
            if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) v = __ldg(src + int64_t(gy) * a.x_sh + gx);
        } else if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
            v = __ldg(src + int64_t(gy) * a.x_sh + gx);
        }
        tin[i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < IH * BL_TW; i += BL_THREADS) {
        const int ly = i / BL_TW, lx = i - ly * BL_TW;
        const float* p = tin + ly * IW + lx;
        float acc = 0.f;
        for (int t = 0; t < k; ++t) acc = fmaf(a.taps[t], p[t], acc);
        tmp[i] = acc;
    }
    __syncthreads();
    float* dst = a.y + int64_t(n) * a.H * a.W;
    for (int i = threadIdx.x; i < BL_TH * BL_TW; i += BL_THREADS) {
        const int ly = i / BL_TW, lx = i - ly * BL_TW;
        const int gy = y0 + ly, gx = x0 + lx;
        if (gy < a.H && gx < a.W) {
            const float* p = tmp + ly * BL_TW + lx;
            float acc = 0.f;
            for (int t = 0; t < k; ++t) acc = fmaf(a.taps[t], p[t * BL_TW], acc);
            dst[int64_t(gy) * a.W + gx] = acc;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Zero-border fast path (k = 3, 5, 7; 16-byte aligned rows): persistent CTAs, a ring of
// TMA-staged halo tiles (out-of-bounds box elements arrive as zeros == the zero padding), one
// warp per 8-row strip, one lane per 4 adjacent columns.  A lane walks down its strip keeping
// the last K horizontally filtered rows in registers; every output row costs one LDS.128 (+2R
// scalar neighbours) and one coalesced STG.128 — x is read from HBM once, y written once.
// ---------------------------------------------------------------------------------------------
#ifndef WM_BT_TH
#define WM_BT_TH 64
#define WM_BT_STAGES 3
#define WM_BT_MINB 2
#endif
constexpr int BT_TW = 128, BT_TH = WM_BT_TH, BT_HALO = 4, BT_BW = BT_TW + 2 * BT_HALO, BT_THREADS = 256, BT_ROWS = BT_TH / (BT_THREADS / 32);

struct BlurTArgs {
    void* y; int N, H, W, tiles_x, tiles_y; int64_t total;
    float taps[8];
    StoreEp ep;
    RaggedSrc rag;      // RAGGED instantiation only: the source planes (no tensor map for rows that are not 16-byte aligned)
};

template <int K> constexpr int bt_stages() { return K <= 5 ? WM_BT_STAGES : 2; }
template <int K> constexpr int bt_stage_floats() { return ((BT_BW * (BT_TH + K - 1) + 31) / 32) * 32; }
// stage stride in elements: whole 128-byte lines for 4-byte and for 2-byte elements
// a TMA box must START on a 16-byte boundary of its row (measured: a 2-byte box at an 8-byte offset is an illegal
// instruction), so the halo of a 2-byte tile is 8 elements
template <int DT> constexpr int bt_halo() { return DT == WM_DT_F32 ? BT_HALO : 8; }
template <int DT> constexpr int bt_bw() { return BT_TW + 2 * bt_halo<DT>(); }
template <int K, int DT> constexpr int bt_stage_elems() { return DT == WM_DT_F32 ? bt_stage_floats<K>() : ((bt_bw<DT>() * (BT_TH + K - 1) + 63) / 64) * 64; }

// RAGGED = rows not 16-byte aligned (W % 4 != 0): the ring is filled by cp.async instead of TMA (stage_box_cpasync) and
// the output leaves by scalar stores; everything between is the same code.
// IDT / ODT = element type of the source planes (the ring holds them as they are; widened on the way to registers)
// and of the result (float32 unless the caller wants the gradient in the autocast type).  Typed instantiations are
// TMA-only and take no store epilogue.
template <int K, bool RAGGED, int IDT = WM_DT_F32, int ODT = WM_DT_F32>
__global__ void __launch_bounds__(BT_THREADS, WM_BT_MINB) gaussblur_tma_kernel(const __grid_constant__ CUtensorMap tmap, const BlurTArgs a) {
    constexpr int R = K / 2, BH = BT_TH + 2 * R, S = bt_stages<K>(), STRIDE = bt_stage_elems<K, IDT>(), ES = tile_elem_size<IDT>();
    constexpr bool PLAIN = IDT == WM_DT_F32 && ODT == WM_DT_F32;
    constexpr int HALO = bt_halo<IDT>(), BW = bt_bw<IDT>();            // staged columns of a tile row (elements)
    static_assert(PLAIN || !RAGGED, "typed planes need 16-byte rows");
    extern __shared__ __align__(128) unsigned char bufs[];
    __shared__ uint64_t full[S];
    const int tid = threadIdx.x;
    if (!RAGGED && tid == 0) {
        tma_prefetch_desc(&tmap);
#pragma unroll
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int per_plane = a.tiles_x * a.tiles_y;
    auto issue = [&](int64_t t, int s) {
        const int n = int(t / per_plane), rem = int(t - int64_t(n) * per_plane);
        const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
        if (RAGGED) {       // every thread; an empty group keeps the per-thread group count in step with the ring
            if (t < a.total) stage_box_cpasync<BT_THREADS>(reinterpret_cast<float*>(bufs) + s * STRIDE, a.rag, n, a.H, a.W, tx * BT_TW - HALO, ty * BT_TH - R, BW, BH);
            else asm volatile("cp.async.commit_group;" ::: "memory");
        } else {
            mbar_expect_tx(&full[s], BW * BH * ES);
            tma_load_3d(bufs + size_t(s) * STRIDE * ES, &tmap, tx * BT_TW - HALO, ty * BT_TH - R, n, &full[s]);
        }
    };
    if (RAGGED || tid == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int64_t t = int64_t(blockIdx.x) + int64_t(s) * gridDim.x;
            if (RAGGED || t < a.total) issue(t, s);
        }
    }
    float w[K];
#pragma unroll
    for (int j = 0; j < K; ++j) w[j] = a.taps[j];
    const int cg = tid & 31, strip = tid >> 5;
    int it = 0;
    for (int64_t t = blockIdx.x; t < a.total; t += gridDim.x, ++it) {
        const int s = it % S;
        if (RAGGED) { cpasync_wait<S - 1>(); __syncthreads(); }      // this thread's copies of tile `it` landed; then everyone's
        else mbar_wait(&full[s], (it / S) & 1);
        const int n = int(t / per_plane), rem = int(t - int64_t(n) * per_plane);
        const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
        const int gx = tx * BT_TW + 4 * cg, gy0 = ty * BT_TH + strip * BT_ROWS;
        const unsigned char* stage = bufs + size_t(s) * STRIDE * ES;
        const int col = (strip * BT_ROWS) * BW + HALO + 4 * cg;            // element index of this lane's first column
        auto hpass = [&](int row, float (&o)[4]) {
            const int p = col + row * BW;
            float win[4 + 2 * R];
            const float4 c = tile_ld4<IDT>(stage, p);
            win[R] = c.x; win[R + 1] = c.y; win[R + 2] = c.z; win[R + 3] = c.w;
#pragma unroll
            for (int i = 0; i < R; ++i) { win[R - 1 - i] = tile_ld1<IDT>(stage, p - 1 - i); win[R + 4 + i] = tile_ld1<IDT>(stage, p + 4 + i); }
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                float acc = w[0] * win[c4];
#pragma unroll
                for (int j = 1; j < K; ++j) acc = fmaf(w[j], win[c4 + j], acc);
                o[c4] = acc;
            }
        };
        float h[K][4];
#pragma unroll
        for (int j = 0; j < K - 1; ++j) hpass(j, h[j]);
        const int64_t doff = (int64_t(n) * a.H + gy0) * a.W + gx;               // element offset of the lane's first output
        float* dst = reinterpret_cast<float*>(a.y) + doff;                        // (float32 results)
        const bool col_ok = gx < a.W;
#pragma unroll
        for (int r = 0; r < BT_ROWS; ++r) {
            hpass(r + K - 1, h[(r + K - 1) % K]);
            float4 o;
            float* op = &o.x;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                float acc = w[0] * h[r % K][c4];
#pragma unroll
                for (int j = 1; j < K; ++j) acc = fmaf(w[j], h[(r + j) % K][c4], acc);
                op[c4] = acc;
            }
            if (col_ok && gy0 + r < a.H) {
                if (PLAIN && a.ep.x)      // x at the output position is the centre of the staged tile row
                    o = a.ep.from_input ? ep_apply4v(o, tile_ld4<IDT>(stage, col + (r + R) * BW), a.ep)
                                        : ep_apply4(o, a.ep.x + doff + int64_t(r) * a.W, a.ep);
                if (RAGGED) st4_ragged(dst + int64_t(r) * a.W, o, gx, a.W);
                else stg4_typed<ODT>(a.y, doff + int64_t(r) * a.W, o);
            }
        }
        __syncthreads();                      // every lane is done with stage s
        if (RAGGED || tid == 0) {
            const int64_t t2 = t + int64_t(S) * gridDim.x;
            if (RAGGED || t2 < a.total) issue(t2, s);
        }
    }
}

template <int K, bool RAGGED, int IDT = WM_DT_F32, int ODT = WM_DT_F32>
static int launch_blur_tma(const void* x, int64_t x_sp, int64_t x_sh, void* y, int N, int H, int W,
                           const float* taps_host, const wm_store_epilogue* ep, cudaStream_t st) {
    constexpr int R = K / 2, S = bt_stages<K>();
    CUtensorMap tm{};
    if (!RAGGED)
        if (int rc = tmap_planes(&tm, tile_tmap_type<IDT>(), tile_elem_size<IDT>(), x, N, H, W, x_sp, x_sh, bt_bw<IDT>(), BT_TH + 2 * R)) {
            set_error("wm_gaussblur: cuTensorMapEncodeTiled failed (%d)", rc);
            return WM_E_ARG;
        }
    BlurTArgs a{};
    a.rag = RaggedSrc{reinterpret_cast<const float*>(x), x_sp, x_sh};
    a.y = y; a.N = N; a.H = H; a.W = W;
    a.tiles_x = (W + BT_TW - 1) / BT_TW; a.tiles_y = (H + BT_TH - 1) / BT_TH;
    a.total = int64_t(N) * a.tiles_x * a.tiles_y;
    for (int i = 0; i < K; ++i) a.taps[i] = taps_host[i];
    a.ep = make_store_ep(ep);
    a.ep.from_input = a.ep.x == x && x_sh == W && x_sp == int64_t(H) * W;      // (a typed call takes no epilogue: ep.x is null)
    const size_t smem = size_t(tile_elem_size<IDT>()) * size_t(S) * bt_stage_elems<K, IDT>();
    cudaError_t e = cudaFuncSetAttribute(gaussblur_tma_kernel<K, RAGGED, IDT, ODT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "wm_gaussblur");
    const int64_t cap = int64_t(sm_count()) * WM_BT_MINB;
    const unsigned grid = (unsigned)(a.total < cap ? a.total : cap);
    gaussblur_tma_kernel<K, RAGGED, IDT, ODT><<<grid, BT_THREADS, smem, st>>>(tm, a);
    WM_LAUNCH_CHECK("wm_gaussblur(tma)");
    return WM_OK;
}

// adjoint of the reflect-border blur, direct gather (GF is never instantiated by the
// reference's trainers; this path favours clarity over speed)
__global__ void __launch_bounds__(256) gaussblur_reflect_adjoint_kernel(const BlurArgs a) {
    const int64_t total = int64_t(a.N) * a.H * a.W;
    const int r = a.r;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int w = int(i % a.W), h = int((i / a.W) % a.H), n = int(i / (int64_t(a.W) * a.H));
        const float* g = a.x + int64_t(n) * a.x_sp;
        // positions of the zero-extended response that fold onto (h, w)
        int ys[3], xs[3], ny = 0, nx = 0;
        ys[ny++] = h; if (h >= 1 && h <= r) ys[ny++] = -h; if (h <= a.H - 2 && h >= a.H - 1 - r) ys[ny++] = 2 * (a.H - 1) - h;
        xs[nx++] = w; if (w >= 1 && w <= r) xs[nx++] = -w; if (w <= a.W - 2 && w >= a.W - 1 - r) xs[nx++] = 2 * (a.W - 1) - w;
        float acc = 0.f;
        for (int iy = 0; iy < ny; ++iy)
            for (int ix = 0; ix < nx; ++ix)
                for (int dy = -r; dy <= r; ++dy) {
                    const int sy = ys[iy] + dy;
                    if (sy < 0 || sy >= a.H) continue;
                    float row = 0.f;
                    for (int dx = -r; dx <= r; ++dx) {
                        const int sx = xs[ix] + dx;
                        if (sx >= 0 && sx < a.W) row = fmaf(a.taps[dx + r], __ldg(g + int64_t(sy) * a.x_sh + sx), row);
                    }
                    acc = fmaf(a.taps[dy + r], row, acc);
                }
        a.y[i] = acc;
    }
}

}  // namespace wm

using namespace wm;

extern "C" int wm_gaussblur(const float* x, int64_t x_sp, int64_t x_sh, float* y, int N, int H, int W,
                            const float* taps_host, int k, int border, int adjoint,
                            const wm_store_epilogue* ep, void* stream) {
    if (N == 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(x && y && taps_host, WM_E_NULL, "wm_gaussblur: null pointer");
    WM_REQUIRE(k >= 1 && k <= BL_MAXK && (k & 1), WM_E_ARG, "wm_gaussblur: kernel size must be odd and <= %d (got %d)", BL_MAXK, k);
    WM_REQUIRE(border == 0 || border == 1, WM_E_ARG, "wm_gaussblur: border must be 0 (zero) or 1 (reflect)");
    WM_REQUIRE(N >= 0 && H > 0 && W > 0, WM_E_SHAPE, "wm_gaussblur: bad shape N=%d H=%d W=%d", N, H, W);
    const int r = (k - 1) / 2;
    WM_REQUIRE(border == 0 || (r < H && r < W), WM_E_SHAPE, "wm_gaussblur: reflect border needs k/2 < H, W");
    if (N == 0) return WM_OK;
    BlurArgs a{};
    a.x = x; a.x_sp = x_sp; a.x_sh = x_sh; a.y = y; a.N = N; a.H = H; a.W = W; a.k = k; a.r = r; a.border = border;
    for (int i = 0; i < k; ++i) a.taps[i] = taps_host[i];
    cudaStream_t st = (cudaStream_t)stream;
    if (border == 1 && adjoint) {
        WM_EP_REJECT(ep, "wm_gaussblur (reflect adjoint)");
        const int64_t total = int64_t(N) * H * W;
        const int64_t want = (total + 255) / 256, cap = int64_t(sm_count()) * 16;
        gaussblur_reflect_adjoint_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(a);
        WM_LAUNCH_CHECK("wm_gaussblur(reflect adjoint)");
        return WM_OK;
    }
    if (border == 0 && (k == 3 || k == 5 || k == 7) && W % 4 == 0 && aligned(y, 16) && tmap_ok(x, x_sp, x_sh, 4)) {
        WM_EP_CHECK(ep, "wm_gaussblur");
        if (k == 3) return launch_blur_tma<3, false>(x, x_sp, x_sh, y, N, H, W, taps_host, ep, st);
        if (k == 5) return launch_blur_tma<5, false>(x, x_sp, x_sh, y, N, H, W, taps_host, ep, st);
        return launch_blur_tma<7, false>(x, x_sp, x_sh, y, N, H, W, taps_host, ep, st);
    }
    if (border == 0 && (k == 3 || k == 5 || k == 7) && !wants_store_ep(ep) && aligned(x, 4) && aligned(y, 4)) {
        // rows that are not 16-byte aligned (W % 4 != 0, odd strides): same ring and stencil, fed by cp.async
        if (k == 3) return launch_blur_tma<3, true>(x, x_sp, x_sh, y, N, H, W, taps_host, nullptr, st);
        if (k == 5) return launch_blur_tma<5, true>(x, x_sp, x_sh, y, N, H, W, taps_host, nullptr, st);
        return launch_blur_tma<7, true>(x, x_sp, x_sh, y, N, H, W, taps_host, nullptr, st);
    }
    WM_EP_REJECT(ep, "wm_gaussblur (generic path)");
    const int IW = BL_TW + 2 * r, IH = BL_TH + 2 * r;
    const size_t smem = sizeof(float) * (size_t(IH) * IW + size_t(IH) * BL_TW);
    cudaError_t e = cudaFuncSetAttribute(gaussblur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "wm_gaussblur");
    const int tiles = ((W + BL_TW - 1) / BL_TW) * ((H + BL_TH - 1) / BL_TH);
    WM_REQUIRE(N <= 65535, WM_E_SHAPE, "wm_gaussblur: at most 65535 planes per call (got %d)", N);
    gaussblur_kernel<<<dim3(tiles, N), BL_THREADS, smem, st>>>(a);
    WM_LAUNCH_CHECK("wm_gaussblur");
    return WM_OK;
}

// Typed planes (include/wm_attack.h): float16 / bfloat16 source read by the ring as it is, or the result stored in
// that type.  Zero border, k = 3 / 5 / 7, rows on 16-byte boundaries - any other case is the caller's to convert.
template <int IDT, int ODT>
static int blur_typed(const void* x, int64_t x_sp, int64_t x_sh, void* y, int N, int H, int W, const float* taps, int k, cudaStream_t st) {
    if (k == 3) return launch_blur_tma<3, false, IDT, ODT>(x, x_sp, x_sh, y, N, H, W, taps, nullptr, st);
    if (k == 5) return launch_blur_tma<5, false, IDT, ODT>(x, x_sp, x_sh, y, N, H, W, taps, nullptr, st);
    return launch_blur_tma<7, false, IDT, ODT>(x, x_sp, x_sh, y, N, H, W, taps, nullptr, st);
}

extern "C" int wm_gaussblur_typed(const void* x, int x_dtype, int64_t x_sp, int64_t x_sh, void* y, int y_dtype, int N, int H, int W,
                                  const float* taps_host, int k, void* stream) {
    if (N == 0) return WM_OK;
    WM_REQUIRE(x && y && taps_host, WM_E_NULL, "wm_gaussblur_typed: null pointer");
    WM_REQUIRE(dtype_ok(x_dtype) && dtype_ok(y_dtype), WM_E_ARG, "wm_gaussblur_typed: unknown element type");
    WM_REQUIRE(x_dtype == WM_DT_F32 || y_dtype == WM_DT_F32, WM_E_ARG, "wm_gaussblur_typed: one side is float32 (typed source OR typed result)");
    WM_REQUIRE(k == 3 || k == 5 || k == 7, WM_E_ARG, "wm_gaussblur_typed: kernel size 3, 5 or 7 (got %d)", k);
    WM_REQUIRE(N > 0 && H > 0 && W > 0, WM_E_SHAPE, "wm_gaussblur_typed: bad shape N=%d H=%d W=%d", N, H, W);
    const size_t xs = dtype_size(x_dtype), ys = dtype_size(y_dtype);
    WM_REQUIRE((W * ys) % (4 * ys) == 0 && W % 4 == 0 && aligned(y, 4 * ys) && tmap_ok(x, x_sp, x_sh, xs), WM_E_ALIGN,
               "wm_gaussblur_typed: rows must start on 16-byte boundaries (W %% %d == 0, aligned planes); convert to float32 otherwise",
               int(16 / xs));
    cudaStream_t st = (cudaStream_t)stream;
    if (x_dtype == WM_DT_F16) return blur_typed<WM_DT_F16, WM_DT_F32>(x, x_sp, x_sh, y, N, H, W, taps_host, k, st);
    if (x_dtype == WM_DT_BF16) return blur_typed<WM_DT_BF16, WM_DT_F32>(x, x_sp, x_sh, y, N, H, W, taps_host, k, st);
    if (y_dtype == WM_DT_F16) return blur_typed<WM_DT_F32, WM_DT_F16>(x, x_sp, x_sh, y, N, H, W, taps_host, k, st);
    if (y_dtype == WM_DT_BF16) return blur_typed<WM_DT_F32, WM_DT_BF16>(x, x_sp, x_sh, y, N, H, W, taps_host, k, st);
    return blur_typed<WM_DT_F32, WM_DT_F32>(x, x_sp, x_sh, y, N, H, W, taps_host, k, st);
}
