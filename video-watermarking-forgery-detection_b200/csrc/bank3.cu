// K-way attack bank, shared read: the 3x3-neighbourhood attacks of a bank — GaussianBlur(k=3),
// MiddleBlur(3), Gaussian noise, Identity — computed from ONE staged tile of the input, each finished by the
// post-attack epilogue (clamp, straight-through, Quantization) and written straight into its slice of the
// [K*B, C, H, W] batch.
//
// Replaces the trainers' K-way attack (models/IRNp_model.py:609-680; per-frame 5-way loop of
// models/IRNcrop_model.py:357-370), where every attack re-reads the batch, five elementwise passes follow each
// attack and a torch.cat copies everything once more.  With one kernel per attack (round 1) a 4-member group costs
// 4 x 24 = 96 B/px; here it costs 12 (read x once) + 4 x 12 (one store per member) = 60 B/px.
//
// Same tile geometry and register ring as median3_tma_kernel (median.cu): persistent CTAs, 3-stage TMA ring of
// 136 x 66 halo tiles (zero fill = the zero padding of both filters), one warp per 8-row strip, one lane per 4
// columns.  Every member's arithmetic is the expression sequence of its stand-alone kernel (blur.cu hpass / vertical
// fmaf chain, median.cu min/max network, philox.cuh Philox + Box-Muller, elementwise.cu clamp, wm_common.cuh ep_apply), so each
// slice is BIT-IDENTICAL to "stand-alone kernel, then the stand-alone epilogue kernel".
#include "philox.cuh"
#include "tma.cuh"
#include "wm_common.cuh"

namespace wm {

// 6 rows per warp strip = two trips of a 3-row body (ring slots are compile-time inside it): the fully unrolled
// 8-row version was 91 KB of SASS and ran from instruction fetch (IPC 0.36)
constexpr int B3_TW = 128, B3_TH = 48, B3_HALO = 4, B3_BW = B3_TW + 2 * B3_HALO, B3_BH = B3_TH + 2,
              B3_THREADS = 256, B3_ROWS = 6, B3_STAGES = 3, B3_STRIDE = ((B3_BW * B3_BH + 31) / 32) * 32;
static_assert(B3_ROWS % 3 == 0 && B3_TH == 8 * B3_ROWS, "strip = whole trips of the 3-row body, 8 warps per tile");

struct Bank3Args {
    float* y_blur; float* y_median; float* y_noise; float* y_identity;
    float taps[3];
    float mean, std; int noise_clamp; unsigned long long seed, offset;
    StoreEp ep;                     // x = the input itself (taken from the staged tile)
    int N, H, W, tiles_x, tiles_y; int64_t total;
};

__device__ __forceinline__ float b3_mid3(float a, float b, float c, float lo, float hi) {
    return __int_as_float(__float_as_int(a) ^ __float_as_int(b) ^ __float_as_int(c) ^ __float_as_int(lo) ^ __float_as_int(hi));
}
__device__ __forceinline__ float b3_med3(float a, float b, float c) { return b3_mid3(a, b, c, fmin3(a, b, c), fmax3(a, b, c)); }

__global__ void __launch_bounds__(B3_THREADS, 2) bank3_kernel(const __grid_constant__ CUtensorMap tmap, const Bank3Args a) {
    extern __shared__ __align__(128) float bufs[];
    __shared__ uint64_t full[B3_STAGES];
    const int tid = threadIdx.x;
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
#pragma unroll
        for (int s = 0; s < B3_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int per_plane = a.tiles_x * a.tiles_y;
    auto issue = [&](int64_t t, int s) {
        const int n = int(t / per_plane), rem = int(t - int64_t(n) * per_plane);
        const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
        mbar_expect_tx(&full[s], B3_BW * B3_BH * sizeof(float));
        tma_load_3d(bufs + s * B3_STRIDE, &tmap, tx * B3_TW - B3_HALO, ty * B3_TH - 1, n, &full[s]);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < B3_STAGES; ++s) {
            const int64_t t = int64_t(blockIdx.x) + int64_t(s) * gridDim.x;
            if (t < a.total) issue(t, s);
        }
    }
    const bool do_blur = a.y_blur != nullptr, do_med = a.y_median != nullptr, do_noise = a.y_noise != nullptr,
               do_id = a.y_identity != nullptr;
    uint64_t seed = a.seed, offset = a.offset;
    if (do_noise) resolve_rng(seed, offset);              // device-resident generator state (CUDA-graph capture)
    const Philox ph(seed);
    const float w0 = a.taps[0], w1 = a.taps[1], w2 = a.taps[2];
    const int cg = tid & 31, strip = tid >> 5;
    int it = 0;
    for (int64_t t = blockIdx.x; t < a.total; t += gridDim.x, ++it) {
        const int s = it % B3_STAGES;
        mbar_wait(&full[s], (it / B3_STAGES) & 1);
        const int n = int(t / per_plane), rem = int(t - int64_t(n) * per_plane);
        const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
        const int gx = tx * B3_TW + 4 * cg, gy0 = ty * B3_TH + strip * B3_ROWS;
        const float* col = bufs + s * B3_STRIDE + (strip * B3_ROWS) * B3_BW + B3_HALO + 4 * cg;
        float raw[3][6], lo[3][4], mi[3][4], hi[3][4], hb[3][4];
        auto load_row = [&](int row, int slot) {
            const float* p = col + row * B3_BW;
            const float4 c = *reinterpret_cast<const float4*>(p);
            raw[slot][0] = p[-1]; raw[slot][1] = c.x; raw[slot][2] = c.y; raw[slot][3] = c.z; raw[slot][4] = c.w;
            raw[slot][5] = p[4];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float u = raw[slot][c4], v = raw[slot][c4 + 1], w = raw[slot][c4 + 2];
                if (do_med) {
                    const float l = fmin3(u, v, w), h = fmax3(u, v, w);
                    lo[slot][c4] = l; hi[slot][c4] = h; mi[slot][c4] = b3_mid3(u, v, w, l, h);
                }
                if (do_blur) {              // blur.cu hpass: w[0]*win[c] then fmaf over the remaining taps
                    float acc = w0 * u;
                    acc = fmaf(w1, v, acc);
                    acc = fmaf(w2, w, acc);
                    hb[slot][c4] = acc;
                }
            }
        };
        load_row(0, 0);
        load_row(1, 1);
        const bool col_ok = gx < a.W;
        const int64_t obase = (int64_t(n) * a.H + gy0) * a.W + gx;
#pragma unroll 1
        for (int r0 = 0; r0 < B3_ROWS; r0 += 3)
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int r = r0 + u;                                     // r % 3 == u: ring slots stay compile-time
            load_row(r + 2, (u + 2) % 3);
            const bool ok = col_ok && gy0 + r < a.H;
            const float* xc = raw[(u + 1) % 3];                       // centre row of the window: x itself
            const float4 xv = make_float4(xc[1], xc[2], xc[3], xc[4]);
            const int64_t o = obase + int64_t(r) * a.W;
            if (do_med) {
                float4 m;
                float* mp = &m.x;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4)
                    mp[c4] = b3_med3(fmax3(lo[0][c4], lo[1][c4], lo[2][c4]), b3_med3(mi[0][c4], mi[1][c4], mi[2][c4]),
                                     fmin3(hi[0][c4], hi[1][c4], hi[2][c4]));
                if (ok) stg128(a.y_median + o, ep_apply4v(m, xv, a.ep));
            }
            if (do_blur) {
                float4 b;
                float* bp = &b.x;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {                      // blur.cu vertical chain over window rows r, r+1, r+2
                    float acc = w0 * hb[u % 3][c4];
                    acc = fmaf(w1, hb[(u + 1) % 3][c4], acc);
                    acc = fmaf(w2, hb[(u + 2) % 3][c4], acc);
                    bp[c4] = acc;
                }
                if (ok) stg128(a.y_blur + o, ep_apply4v(b, xv, a.ep));
            }
            if (do_noise && ok) {
                float4 nz = normal4(ph, (uint64_t)(o >> 2) + offset);
                nz.x = fmaf(nz.x, a.std, a.mean); nz.y = fmaf(nz.y, a.std, a.mean);
                nz.z = fmaf(nz.z, a.std, a.mean); nz.w = fmaf(nz.w, a.std, a.mean);
                float4 v = make_float4(xv.x + nz.x, xv.y + nz.y, xv.z + nz.z, xv.w + nz.w);
                if (a.noise_clamp) v = clamp01_nan4(v);
                v = ep_apply4v(v, xv, a.ep);
                stg128(a.y_noise + o, v);
            }
            if (do_id && ok) stg128(a.y_identity + o, ep_apply4v(xv, xv, a.ep));
        }
        __syncthreads();
        if (tid == 0) {
            const int64_t t2 = t + int64_t(B3_STAGES) * gridDim.x;
            if (t2 < a.total) issue(t2, s);
        }
    }
}

}  // namespace wm

using namespace wm;

extern "C" int wm_bank3_ok(int N, int H, int W) {
    return (N > 0 && N <= 65535 && H > 0 && W > 0 && W % 4 == 0 && tmap_encoder() != nullptr) ? 1 : 0;
}

extern "C" int wm_bank3_fwd(const float* x, int64_t x_sp, int64_t x_sh, int N, int H, int W,
                            const wm_bank3_desc* d, void* stream) {
    if (N == 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(x && d, WM_E_NULL, "wm_bank3_fwd: null pointer");
    WM_REQUIRE(d->y_blur || d->y_median || d->y_noise || d->y_identity, WM_E_ARG, "wm_bank3_fwd: no member selected");
    WM_REQUIRE(wm_bank3_ok(N, H, W), WM_E_SHAPE, "wm_bank3_fwd: needs W %% 4 == 0 and 0 < N <= 65535 (got N=%d H=%d W=%d)", N, H, W);
    WM_REQUIRE(tmap_ok(x, x_sp, x_sh, 4), WM_E_ALIGN, "wm_bank3_fwd: x must be 16-byte aligned with strides multiples of 4 elements");
    const float* outs[4] = {d->y_blur, d->y_median, d->y_noise, d->y_identity};
    for (const float* p : outs) WM_REQUIRE(p == nullptr || aligned(p, 16), WM_E_ALIGN, "wm_bank3_fwd: outputs must be 16-byte aligned");
    CUtensorMap tm;
    if (int rc = tmap_planes(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, N, H, W, x_sp, x_sh, B3_BW, B3_BH)) {
        set_error("wm_bank3_fwd: cuTensorMapEncodeTiled failed (%d)", rc);
        return WM_E_ARG;
    }
    Bank3Args a{};
    a.y_blur = d->y_blur; a.y_median = d->y_median; a.y_noise = d->y_noise; a.y_identity = d->y_identity;
    for (int i = 0; i < 3; ++i) a.taps[i] = d->blur_taps[i];
    a.mean = d->noise_mean; a.std = d->noise_std; a.noise_clamp = d->noise_clamp; a.seed = d->seed; a.offset = d->offset;
    a.ep = StoreEp{x, d->clamp01, d->quantize, 1};
    a.N = N; a.H = H; a.W = W; a.tiles_x = (W + B3_TW - 1) / B3_TW; a.tiles_y = (H + B3_TH - 1) / B3_TH;
    a.total = int64_t(N) * a.tiles_x * a.tiles_y;
    const size_t smem = sizeof(float) * size_t(B3_STAGES) * B3_STRIDE;
    cudaError_t e = cudaFuncSetAttribute(bank3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "wm_bank3_fwd");
    const int64_t cap = int64_t(sm_count()) * 2;
    bank3_kernel<<<(unsigned)(a.total < cap ? a.total : cap), B3_THREADS, smem, (cudaStream_t)stream>>>(tm, a);
    WM_LAUNCH_CHECK("wm_bank3_fwd");
    return WM_OK;
}
