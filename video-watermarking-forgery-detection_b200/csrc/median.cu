// k x k median filter (k = 3, 5), zero padding, over N = B*C planes.
//
// Replaces MiddleBlur.forward (noise_layers/middle_filter.py:11-13 -> kornia MedianBlur), which
// materialises a one-hot conv2d expansion of k*k times the image (288 / 672 B/px forward) and
// then sorts along it.  Here a CTA stages a halo tile in shared memory; a thread walks down a
// column strip, sorts each k-wide window ROW once (shared by the k vertically adjacent outputs)
// and selects the median from the k sorted rows with a pruned min/max network
// (3x3: FMNMX3 + XOR middle-of-three, 5x5: generated median_net.cuh) — pure selection, so the
// forward is bit-exact.  Optionally it records, per output, the raster position of the FIRST
// window element equal to the median (uint8); the backward is a deterministic gather through it.
#include "median_net.cuh"
#include "wm_common.cuh"

namespace wm {

constexpr int MD_TW = 128, MD_TH = 32, MD_THREADS = 256, MD_STRIP = MD_TH * MD_TW / MD_THREADS;  // 16 rows/thread

struct MedArgs {
    const float* x; int64_t x_sp, x_sh;
    float* y; uint8_t* idx; int N, H, W;
};

__device__ __forceinline__ float mid3(float a, float b, float c, float lo, float hi) {
    // the element that is neither the min nor the max: XOR of the five bit patterns
    return __int_as_float(__float_as_int(a) ^ __float_as_int(b) ^ __float_as_int(c) ^
                          __float_as_int(lo) ^ __float_as_int(hi));
}
__device__ __forceinline__ void sort3(float& a, float& b, float& c) {
    const float lo = fmin3(a, b, c), hi = fmax3(a, b, c);
    b = mid3(a, b, c, lo, hi); a = lo; c = hi;
}
__device__ __forceinline__ float med3(float a, float b, float c) {
    return mid3(a, b, c, fmin3(a, b, c), fmax3(a, b, c));
}
#define MD_CE(a, b) { const float lo__ = fminf(a, b); b = fmaxf(a, b); a = lo__; }
__device__ __forceinline__ void sort5(float (&v)[5]) {
    // optimal 9-comparator network
    MD_CE(v[0], v[1]); MD_CE(v[3], v[4]); MD_CE(v[2], v[4]); MD_CE(v[2], v[3]); MD_CE(v[0], v[3]);
    MD_CE(v[0], v[2]); MD_CE(v[1], v[4]); MD_CE(v[1], v[3]); MD_CE(v[1], v[2]);
}
#undef MD_CE

template <int K, bool WANT_IDX>
__global__ void __launch_bounds__(MD_THREADS) median_fwd_kernel(const MedArgs a) {
    constexpr int R = K / 2, IW = MD_TW + 2 * R, IH = MD_TH + 2 * R;
    __shared__ float tile[IH * IW];
    const int tiles_x = (a.W + MD_TW - 1) / MD_TW;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int n = blockIdx.y;
    const int x0 = tx * MD_TW, y0 = ty * MD_TH;
    const float* src = a.x + int64_t(n) * a.x_sp;
    for (int i = threadIdx.x; i < IH * IW; i += MD_THREADS) {
        const int ly = i / IW, lx = i - ly * IW;
        const int gy = y0 + ly - R, gx = x0 + lx - R;
        tile[i] = (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) ? __ldg(src + int64_t(gy) * a.x_sh + gx) : 0.f;
    }
    __syncthreads();
    const int lx = threadIdx.x % MD_TW, strip = threadIdx.x / MD_TW;
    const int ly0 = strip * MD_STRIP;
    const int gx = x0 + lx;
    float* dst = a.y + int64_t(n) * a.H * a.W;
    uint8_t* dsti = WANT_IDX ? a.idx + int64_t(n) * a.H * a.W : nullptr;

    float rows[K][K];       // ring of the K sorted window rows
#pragma unroll
    for (int j = 0; j < K - 1; ++j) {
#pragma unroll
        for (int c = 0; c < K; ++c) rows[j][c] = tile[(ly0 + j) * IW + lx + c];
        if constexpr (K == 3) sort3(rows[j][0], rows[j][1], rows[j][2]);
        else sort5(rows[j]);
    }
#pragma unroll
    for (int s = 0; s < MD_STRIP; ++s) {
        const int ly = ly0 + s;
        constexpr int KM1 = K - 1;
        const int slot = (s + KM1) % K;     // compile-time after unrolling: replaces the oldest row
#pragma unroll
        for (int c = 0; c < K; ++c) rows[slot][c] = tile[(ly + K - 1) * IW + lx + c];
        float med;
        if constexpr (K == 3) {
            sort3(rows[slot][0], rows[slot][1], rows[slot][2]);
            med = med3(fmax3(rows[0][0], rows[1][0], rows[2][0]),
                       med3(rows[0][1], rows[1][1], rows[2][1]),
                       fmin3(rows[0][2], rows[1][2], rows[2][2]));
        } else {
            sort5(rows[slot]);
            float v[25];
#pragma unroll
            for (int j = 0; j < 5; ++j)
#pragma unroll
                for (int c = 0; c < 5; ++c) v[5 * j + c] = rows[j][c];
            med = median25_sorted_groups(v);
        }
        const int gy = y0 + ly;
        if (gy < a.H && gx < a.W) {
            dst[int64_t(gy) * a.W + gx] = med;
            if (WANT_IDX) {
                int pos = 0;
#pragma unroll
                for (int j = K * K - 1; j >= 0; --j)
                    pos = (tile[(ly + j / K) * IW + lx + j % K] == med) ? j : pos;
                dsti[int64_t(gy) * a.W + gx] = (uint8_t)pos;
            }
        }
    }
}

// gx[p] = sum over outputs q with p in window(q) and argmedian(q) == p of gy[q]
template <int K>
__global__ void __launch_bounds__(256) median_bwd_kernel(const float* __restrict__ gy, const uint8_t* __restrict__ idx,
                                                         float* __restrict__ gx, int N, int H, int W) {
    constexpr int R = K / 2;
    const int64_t total = int64_t(N) * H * W;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int w = int(i % W), h = int((i / W) % H);
        const int64_t base = i - (int64_t(h) * W + w);
        float acc = 0.f;
#pragma unroll
        for (int dy = -R; dy <= R; ++dy)
#pragma unroll
            for (int dx = -R; dx <= R; ++dx) {
                const int qh = h + dy, qw = w + dx;     // output whose window contains (h, w)
                if (qh >= 0 && qh < H && qw >= 0 && qw < W) {
                    // (h, w) sits at window position (R - dy, R - dx) of output (qh, qw)
                    const int want = (R - dy) * K + (R - dx);
                    const int64_t q = base + int64_t(qh) * W + qw;
                    if (idx[q] == want) acc += gy[q];
                }
            }
        gx[i] = acc;
    }
}

}  // namespace wm

using namespace wm;

extern "C" int wm_median_fwd(const float* x, int64_t x_sp, int64_t x_sh, float* y, uint8_t* idx,
                             int N, int H, int W, int k, void* stream) {
    WM_REQUIRE(x && y, WM_E_NULL, "wm_median_fwd: null pointer");
    WM_REQUIRE(k == 3 || k == 5, WM_E_ARG, "wm_median_fwd: kernel size must be 3 or 5 (got %d)", k);
    WM_REQUIRE(N >= 0 && N <= 65535 && H > 0 && W > 0, WM_E_SHAPE, "wm_median_fwd: bad shape N=%d H=%d W=%d", N, H, W);
    if (N == 0) return WM_OK;
    MedArgs a{x, x_sp, x_sh, y, idx, N, H, W};
    const int tiles = ((W + MD_TW - 1) / MD_TW) * ((H + MD_TH - 1) / MD_TH);
    dim3 grid(tiles, N);
    cudaStream_t st = (cudaStream_t)stream;
    if (k == 3) { if (idx) median_fwd_kernel<3, true><<<grid, MD_THREADS, 0, st>>>(a); else median_fwd_kernel<3, false><<<grid, MD_THREADS, 0, st>>>(a); }
    else        { if (idx) median_fwd_kernel<5, true><<<grid, MD_THREADS, 0, st>>>(a); else median_fwd_kernel<5, false><<<grid, MD_THREADS, 0, st>>>(a); }
    WM_LAUNCH_CHECK("wm_median_fwd");
    return WM_OK;
}

extern "C" int wm_median_bwd(const float* gy, const uint8_t* idx, float* gx, int N, int H, int W, int k, void* stream) {
    WM_REQUIRE(gy && idx && gx, WM_E_NULL, "wm_median_bwd: null pointer");
    WM_REQUIRE(k == 3 || k == 5, WM_E_ARG, "wm_median_bwd: kernel size must be 3 or 5 (got %d)", k);
    const int64_t total = int64_t(N) * H * W;
    if (total <= 0) return WM_OK;
    const int64_t want = (total + 255) / 256, cap = int64_t(sm_count()) * 32;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (k == 3) median_bwd_kernel<3><<<grid, 256, 0, (cudaStream_t)stream>>>(gy, idx, gx, N, H, W);
    else median_bwd_kernel<5><<<grid, 256, 0, (cudaStream_t)stream>>>(gy, idx, gx, N, H, W);
    WM_LAUNCH_CHECK("wm_median_bwd");
    return WM_OK;
}
