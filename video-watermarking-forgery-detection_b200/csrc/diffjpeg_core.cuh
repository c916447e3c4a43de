// Fused DiffJPEG forward / analytic backward / compress / decompress for sm_100a
// (shared device code; kernels live in diffjpeg_{fwd,bwd,codec}.cu so they compile in parallel).
//
// Reference path replaced: utils/JPEG.py:501-540 (DiffJPEG.forward = decompress_jpeg(compress_jpeg(x)))
// — ~93 materialising torch kernels and 876 B/px of HBM traffic forward+backward — by ONE
// kernel per direction that touches HBM for the compulsory bytes only
// (forward: read x 12 B/px, write y 12 B/px; backward: read x, gy, write gx = 36 B/px).
//
// Mapping (no tensor cores: an 8x8 DCT is not a contraction worth tcgen05):
//   * one THREAD owns one 8x8 luminance block; the 4 threads (lane bits 0 and 4) that own the
//     2x2 luminance blocks of a 16x16 MCU jointly own its Cb and Cr blocks as 4x4 quadrants;
//   * a warp = 8 consecutive MCUs: lanes 0-15 hold the upper block row, 16-31 the lower, so
//     each 256-bit load/store instruction (LDG.E.256 / STG.E.256) of a warp covers two fully
//     used contiguous 512-byte runs — every 32-byte sector moved is used once;
//   * the 2-D DCT of the luminance block streams through a thread-private shared-memory
//     scratch (row pass in registers as rows arrive -> scratch -> 4-column groups -> scratch
//     -> row pass as rows leave), which keeps the register footprint at ~100 instead of the
//     192 a register-resident RGB block would need;
//   * the chroma block uses the divergence-free 4-lane split transform of dct8.cuh
//     (__shfl_xor with lane^1 / lane^16).
#pragma once
#include "dct8.cuh"
#include "wm_common.cuh"

namespace wm {

// T_Y[u][v] = AnnexK[v][u]: the reference stores the luminance table TRANSPOSED
// (utils/JPEG.py:98-104) and multiplies it un-rounded by `factor` (:229, :310).
static __constant__ float cTY[64] = {
    16, 12, 14, 14, 18, 24, 49, 72,
    11, 12, 13, 17, 22, 35, 64, 92,
    10, 14, 16, 22, 37, 55, 78, 95,
    16, 19, 24, 29, 56, 64, 87, 98,
    24, 26, 40, 51, 68, 81, 103, 112,
    40, 58, 57, 87, 109, 104, 121, 100,
    51, 60, 69, 80, 103, 113, 120, 103,
    61, 55, 56, 62, 77, 92, 101, 99};
// utils/JPEG.py:107-110 (symmetric)
static __constant__ float cTC[64] = {
    17, 18, 24, 47, 99, 99, 99, 99,
    18, 21, 26, 66, 99, 99, 99, 99,
    24, 26, 56, 99, 99, 99, 99, 99,
    47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99};

constexpr int DJ_THREADS = 128;

// ---- rounding surrogates (utils/JPEG.py:472-484, utils/JPEG_utils.py:36-41) -----------------
template <int MODE>
__device__ __forceinline__ float round_fwd(float q) {
    if (MODE == WM_ROUND_ONLY_AT_0) return fabsf(q) < 0.5f ? q * q * q : q;
    if (MODE == WM_ROUND_CUBIC) { float r = rintf(q), d = q - r; return fmaf(d * d, d, r); }
    if (MODE == WM_ROUND_HARD) return rintf(q);
    // Fourier: q - (1/pi) sum_{n=1..9} (-1)^{n+1}/n sin(2 pi n q)
    float s = 0.f;
#pragma unroll
    for (int n = 1; n <= 9; ++n) s += ((n & 1) ? 1.f : -1.f) / n * sinpif(2.f * n * q);
    return q - s * 0.318309886183790672f;
}
template <int MODE>
__device__ __forceinline__ float round_grad(float q) {
    if (MODE == WM_ROUND_ONLY_AT_0) return fabsf(q) < 0.5f ? 3.f * q * q : 1.f;
    if (MODE == WM_ROUND_CUBIC) { float d = q - rintf(q); return 3.f * d * d; }
    if (MODE == WM_ROUND_HARD) return 0.f;
    float s = 0.f;
#pragma unroll
    for (int n = 1; n <= 9; ++n) s += ((n & 1) ? 1.f : -1.f) * cospif(2.f * n * q);
    return 1.f - 2.f * s;
}

// ---- colour constants (utils/JPEG.py:125-135, :419-428) --------------------------------------
#define DJ_YR (0.299f * 255.f)
#define DJ_YG (0.587f * 255.f)
#define DJ_YB (0.114f * 255.f)
#define DJ_CBR (-0.168736f * 255.f * 0.25f)
#define DJ_CBG (-0.331264f * 255.f * 0.25f)
#define DJ_CBB (0.5f * 255.f * 0.25f)
#define DJ_CRR (0.5f * 255.f * 0.25f)
#define DJ_CRG (-0.418688f * 255.f * 0.25f)
#define DJ_CRB (-0.081312f * 255.f * 0.25f)
#define DJ_I255 (1.f / 255.f)

struct DJArgs {
    const float* x; int64_t x_sb, x_sc, x_sh;
    const float* gy; int64_t g_sb, g_sc, g_sh;
    float* out;                 // y (fwd) or gx (bwd), dense NCHW
    float* coef_y; float* coef_cb; float* coef_cr;  // compress / decompress
    int B, H, W;
    int mcu_w, mcu_per_img; int64_t n_mcu;
    float factor; const float* factor_ps;
};

struct DJThread {
    bool active; int b, row0, col0, bx, by, mcu_y, mcu_x;
};

__device__ __forceinline__ DJThread dj_locate(const DJArgs& a) {
    DJThread t;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t mcu = warp * 8 + ((lane & 15) >> 1);
    t.bx = lane & 1; t.by = lane >> 4;
    t.active = mcu < a.n_mcu;
    const int64_t m = t.active ? mcu : 0;
    t.b = int(m / a.mcu_per_img);
    const int rem = int(m - int64_t(t.b) * a.mcu_per_img);
    t.mcu_y = rem / a.mcu_w; t.mcu_x = rem - t.mcu_y * a.mcu_w;
    t.row0 = t.mcu_y * 16 + t.by * 8; t.col0 = t.mcu_x * 16 + t.bx * 8;
    return t;
}

// scratch: 16 float4 chunks per thread, chunk c of thread t at [c * DJ_THREADS + t]
// (consecutive lanes -> consecutive 16-byte words: conflict-free LDS.128/STS.128)
__device__ __forceinline__ void scr_store_row(float4* scr, int r, const float (&v)[8]) {
    scr[(2 * r) * DJ_THREADS] = make_float4(v[0], v[1], v[2], v[3]);
    scr[(2 * r + 1) * DJ_THREADS] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void scr_load_row(const float4* scr, int r, float (&v)[8]) {
    float4 a = scr[(2 * r) * DJ_THREADS], b = scr[(2 * r + 1) * DJ_THREADS];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// Phase 1: stream the RGB rows of the thread's 8x8 block: luminance rows are level-shifted,
// row-transformed and parked in scratch; chroma is reduced to the 4x4 quadrant (2x2 mean).
__device__ __forceinline__ void dj_load_block(const DJArgs& a, const DJThread& t, float4* scr,
                                              float (&cb)[4][4], float (&cr)[4][4]) {
    const float* xr = a.x + int64_t(t.b) * a.x_sb + int64_t(t.row0) * a.x_sh + t.col0;
#pragma unroll
    for (int rp = 0; rp < 4; ++rp) {
        f8 R[2], G[2], Bl[2];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const float* p = xr + int64_t(2 * rp + rr) * a.x_sh;
            if (t.active) {
                R[rr] = ldg256_stream(p);
                G[rr] = ldg256_stream(p + a.x_sc);
                Bl[rr] = ldg256_stream(p + 2 * a.x_sc);
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) R[rr].v[c] = G[rr].v[c] = Bl[rr].v[c] = 0.f;
            }
        }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            float yv[8];
#pragma unroll
            for (int c = 0; c < 8; ++c)
                yv[c] = fmaf(R[rr].v[c], DJ_YR, fmaf(G[rr].v[c], DJ_YG, fmaf(Bl[rr].v[c], DJ_YB, -128.f)));
            dct8(yv);
            scr_store_row(scr, 2 * rp + rr, yv);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float sr = (R[0].v[2 * j] + R[0].v[2 * j + 1]) + (R[1].v[2 * j] + R[1].v[2 * j + 1]);
            const float sg = (G[0].v[2 * j] + G[0].v[2 * j + 1]) + (G[1].v[2 * j] + G[1].v[2 * j + 1]);
            const float sb = (Bl[0].v[2 * j] + Bl[0].v[2 * j + 1]) + (Bl[1].v[2 * j] + Bl[1].v[2 * j + 1]);
            cb[rp][j] = fmaf(sr, DJ_CBR, fmaf(sg, DJ_CBG, sb * DJ_CBB));   // (Cb + 128) - 128
            cr[rp][j] = fmaf(sr, DJ_CRR, fmaf(sg, DJ_CRG, sb * DJ_CRB));
        }
    }
}

// Phase 2: luminance column stage on 4-column groups.
//   KEEP_Q : write the rounded quantised coefficient back (compress) instead of the
//            dequantised, column-inverse-transformed value (forward / backward recompute)
//   GRAD   : also park d round / dq in scratch region `dscr`
template <int ROUND, bool KEEP_Q, bool GRAD>
__device__ __forceinline__ void dj_luma_columns(float4* scr, float4* dscr, float f) {
#pragma unroll
    for (int cg = 0; cg < 2; ++cg) {
        float v[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float4 t4 = scr[(2 * r + cg) * DJ_THREADS];
            v[r][0] = t4.x; v[r][1] = t4.y; v[r][2] = t4.z; v[r][3] = t4.w;
        }
        float d[8][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            dct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float tf = cTY[u * 8 + 4 * cg + j] * f;
                const float q = div_by_recip(v[u][j], tf, fast_rcp(tf));
                if (GRAD) d[u][j] = round_grad<ROUND>(q);
                const float rq = round_fwd<ROUND>(q);
                v[u][j] = KEEP_Q ? rq : rq * tf;
            }
            if (!KEEP_Q)
                idct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            scr[(2 * r + cg) * DJ_THREADS] = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
            if (GRAD) dscr[(2 * r + cg) * DJ_THREADS] = make_float4(d[r][0], d[r][1], d[r][2], d[r][3]);
        }
    }
}

// Phase 3: chroma quadrant -> split DCT -> quantise/round/dequantise -> split IDCT.
template <int ROUND, bool KEEP_Q, bool GRAD>
__device__ __forceinline__ void dj_chroma_roundtrip(float (&p)[4][4], float (&d)[4][4],
                                                    const QuadCoef& qx, const QuadCoef& qy,
                                                    int bx, int by, float f) {
    quad_dct_rows(p, qx, 1);
    quad_dct_cols(p, qy, 16);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float tf = cTC[(16 * i + 2 * j) + (8 * by + bx)] * f;   // [u=2i+by][v=2j+bx]
            const float q = div_by_recip(p[i][j], tf, fast_rcp(tf));
            if (GRAD) d[i][j] = round_grad<ROUND>(q);
            const float rq = round_fwd<ROUND>(q);
            p[i][j] = KEEP_Q ? rq : rq * tf;
        }
    if (!KEEP_Q) {
        quad_idct_cols(p, qy, 16);
        quad_idct_rows(p, qx, 1);
    }
}

__device__ __forceinline__ void dj_chroma_terms(const float (&cb)[4], const float (&cr)[4],
                                                float (&tR)[4], float (&tG)[4], float (&tB)[4]) {
    // (Y' + 128 + k * C'') / 255 with the +128 and the /255 folded into the chroma term
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        tR[j] = fmaf(cr[j], 1.402f * DJ_I255, 128.f * DJ_I255);
        tG[j] = fmaf(cb[j], -0.344136f * DJ_I255, fmaf(cr[j], -0.714136f * DJ_I255, 128.f * DJ_I255));
        tB[j] = fmaf(cb[j], 1.772f * DJ_I255, 128.f * DJ_I255);
    }
}

// ---------------------------------------------------------------------------------------------
static inline int dj_check(const float* x, int64_t sb, int64_t sc, int64_t sh, int B, int H, int W, const char* who) {
    WM_REQUIRE(x != nullptr, WM_E_NULL, "%s: null image pointer", who);
    WM_REQUIRE(B >= 0 && H > 0 && W > 0 && H % 16 == 0 && W % 16 == 0, WM_E_SHAPE,
               "%s: H and W must be positive multiples of 16 (got B=%d H=%d W=%d); the reference's "
               "block_merging views require it (utils/JPEG.py:371-376)", who, B, H, W);
    WM_REQUIRE(aligned(x, 32) && sb % 8 == 0 && sc % 8 == 0 && sh % 8 == 0, WM_E_ALIGN,
               "%s: base pointer must be 32-byte aligned and strides multiples of 8 elements "
               "(sb=%lld sc=%lld sh=%lld)", who, (long long)sb, (long long)sc, (long long)sh);
    return WM_OK;
}

static inline DJArgs dj_args(int B, int H, int W, float factor, const float* ps) {
    DJArgs a{};
    a.B = B; a.H = H; a.W = W;
    a.mcu_w = W / 16; a.mcu_per_img = (H / 16) * (W / 16);
    a.n_mcu = int64_t(B) * a.mcu_per_img;
    a.factor = factor; a.factor_ps = ps;
    return a;
}

template <typename K>
static inline int dj_launch(K kernel, const DJArgs& a, size_t smem, cudaStream_t st, const char* who) {
    if (a.n_mcu == 0) return WM_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, who);
    const int64_t warps = (a.n_mcu + 7) / 8;
    const int64_t blocks = (warps * 32 + DJ_THREADS - 1) / DJ_THREADS;
    kernel<<<(unsigned)blocks, DJ_THREADS, smem, st>>>(a);
    WM_LAUNCH_CHECK(who);
    return WM_OK;
}

}  // namespace wm

#define DJ_DISPATCH_ROUND(KERNEL, ...)                                                        \
    switch (rounding) {                                                                       \
        case WM_ROUND_ONLY_AT_0: return dj_launch(KERNEL<WM_ROUND_ONLY_AT_0>, __VA_ARGS__);   \
        case WM_ROUND_CUBIC:     return dj_launch(KERNEL<WM_ROUND_CUBIC>, __VA_ARGS__);       \
        case WM_ROUND_HARD:      return dj_launch(KERNEL<WM_ROUND_HARD>, __VA_ARGS__);        \
        case WM_ROUND_FOURIER:   return dj_launch(KERNEL<WM_ROUND_FOURIER>, __VA_ARGS__);     \
        default: set_error("unknown rounding mode %d", rounding); return WM_E_ARG;            \
    }

