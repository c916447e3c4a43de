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
//     192 a register-resident RGB block would need; the chroma quadrants park there too, so
//     every phase is a ROLLED loop over row pairs / column groups / planes: the whole kernel
//     stays inside the instruction cache (the first, fully unrolled version was 150 KB of SASS
//     and spent half its cycles on instruction fetch — profiles/ncu_r1_diffjpeg_unrolled.txt);
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

constexpr int DJ_THREADS = 128;    // forward / compress / decompress CTA
constexpr int DJB_THREADS = 96;    // backward CTA (48 scratch chunks per thread -> 3 CTAs per SM)

// Thread-private scratch layout, in float4 chunks (chunk c of thread t at [c * NT + t], so
// consecutive lanes touch consecutive 16-byte words: conflict-free LDS.128 / STS.128):
constexpr int SC_Y = 0;      // 16 chunks: luminance block, row r = chunks 2r, 2r+1
constexpr int SC_CB = 16;    //  4 chunks: Cb quadrant rows   (backward: later the Cb cotangent)
constexpr int SC_CR = 20;    //  4 chunks: Cr quadrant rows
constexpr int SC_DY = 24;    // 16 chunks: round'(q) of the luminance block      (backward only)
constexpr int SC_DC = 40;    //  8 chunks: round'(q) of the Cb, Cr quadrants     (backward only)
constexpr int SC_FWD_CHUNKS = 24, SC_BWD_CHUNKS = 48;

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
    const void* x; int x_dt; int64_t x_sb, x_sc, x_sh;   // image in WM_DT_* elements (cast fused into the load)
    const float* gy; int64_t g_sb, g_sc, g_sh;
    float* out;                 // y (fwd) or gx (bwd), dense NCHW; the backward kernels store gx as out_dt elements
    int out_dt;
    float* coef_y; float* coef_cb; float* coef_cr;  // compress / decompress
    // state saved by the forward for the backward (wm_diffjpeg_fwd_save / wm_diffjpeg_bwd_saved):
    float* dY;                  // [B, H, W]        round'(q) of the luminance coefficient at [8i+u, 8j+v]
    float* dC;                  // [B, 2, H/2, W/2] round'(q) of Cb, Cr (per-thread quadrant order)
    unsigned long long* cm;     // [B, H, W/8]      clamp codes of one 8-pixel row: 2 bits x (8 px x RGB)
    StoreEp ep;                 // forward-only kernel: store epilogue (x dense, same layout as out)
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

__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

template <int NT>
__device__ __forceinline__ void scr_store_row(float4* scr, int r, const float (&v)[8]) {
    scr[(SC_Y + 2 * r) * NT] = make_float4(v[0], v[1], v[2], v[3]);
    scr[(SC_Y + 2 * r + 1) * NT] = make_float4(v[4], v[5], v[6], v[7]);
}
template <int NT>
__device__ __forceinline__ void scr_load_row(const float4* scr, int r, float (&v)[8]) {
    const float4 a = scr[(SC_Y + 2 * r) * NT], b = scr[(SC_Y + 2 * r + 1) * NT];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void f4_to(float (&v)[4], const float4 a) { v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; }
__device__ __forceinline__ float4 to_f4(const float (&v)[4]) { return make_float4(v[0], v[1], v[2], v[3]); }

struct RowPair { f8 R[2], G[2], B[2]; };

__device__ __forceinline__ void dj_load_pair(RowPair& p, const void* base, int64_t off, int64_t sh, int64_t sc, bool active,
                                             int dt = WM_DT_F32) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        const int64_t q = off + int64_t(rr) * sh;
        if (active) {
            p.R[rr] = ld8_typed(base, q, dt);
            p.G[rr] = ld8_typed(base, q + sc, dt);
            p.B[rr] = ld8_typed(base, q + 2 * sc, dt);
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) p.R[rr].v[c] = p.G[rr].v[c] = p.B[rr].v[c] = 0.f;
        }
    }
}

// luminance rows are level-shifted, row-transformed and parked; chroma is 2x2-averaged
template <int NT>
__device__ __forceinline__ void dj_consume_pair(const RowPair& p, int rp, float4* scr) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        float yv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
            yv[c] = fmaf(p.R[rr].v[c], DJ_YR, fmaf(p.G[rr].v[c], DJ_YG, fmaf(p.B[rr].v[c], DJ_YB, -128.f)));
        dct8(yv);
        scr_store_row<NT>(scr, 2 * rp + rr, yv);
    }
    float cb[4], cr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float sr = (p.R[0].v[2 * j] + p.R[0].v[2 * j + 1]) + (p.R[1].v[2 * j] + p.R[1].v[2 * j + 1]);
        const float sg = (p.G[0].v[2 * j] + p.G[0].v[2 * j + 1]) + (p.G[1].v[2 * j] + p.G[1].v[2 * j + 1]);
        const float sb = (p.B[0].v[2 * j] + p.B[0].v[2 * j + 1]) + (p.B[1].v[2 * j] + p.B[1].v[2 * j + 1]);
        cb[j] = fmaf(sr, DJ_CBR, fmaf(sg, DJ_CBG, sb * DJ_CBB));   // (Cb + 128) - 128
        cr[j] = fmaf(sr, DJ_CRR, fmaf(sg, DJ_CRG, sb * DJ_CRB));
    }
    scr[(SC_CB + rp) * NT] = to_f4(cb);
    scr[(SC_CR + rp) * NT] = to_f4(cr);
}

// Phase 1: stream the 8 RGB rows of the thread's block, two rows at a time, always one pair
// of loads (6 x LDG.E.256) in flight ahead of the arithmetic.  Rolled (2 iterations) to keep
// the kernel inside the instruction cache.
// TYPED = false: the image is float32 and the element-type switch folds away at compile time (the float32 kernels
// are exactly the pre-typed ones); TYPED = true instantiations read a.x_dt / store a.out_dt elements.
template <int NT, bool TYPED = false>
__device__ __forceinline__ void dj_load_block(const DJArgs& a, const DJThread& t, float4* scr) {
    const int64_t xr = int64_t(t.b) * a.x_sb + int64_t(t.row0) * a.x_sh + t.col0;
    const int xdt = TYPED ? a.x_dt : WM_DT_F32;
    RowPair A, B;
    dj_load_pair(A, a.x, xr, a.x_sh, a.x_sc, t.active, xdt);
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
        dj_load_pair(B, a.x, xr + int64_t(4 * it + 2) * a.x_sh, a.x_sh, a.x_sc, t.active, xdt);
        dj_consume_pair<NT>(A, 2 * it, scr);
        if (it == 0) dj_load_pair(A, a.x, xr + int64_t(4) * a.x_sh, a.x_sh, a.x_sc, t.active, xdt);
        dj_consume_pair<NT>(B, 2 * it + 1, scr);
    }
}

// Phase 2: luminance column stage on 4-column groups (rolled over the two groups).
//   KEEP_Q : write the rounded quantised coefficient back (compress) instead of the
//            dequantised, column-inverse-transformed value (forward / backward recompute)
//   GRAD   : also park d round / dq in the SC_DY chunks
template <int ROUND, bool KEEP_Q, bool GRAD, int NT>
__device__ __forceinline__ void dj_luma_columns(float4* scr, float f) {
#pragma unroll 1
    for (int cg = 0; cg < 2; ++cg) {
        float v[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r) f4_to(v[r], scr[(SC_Y + 2 * r + cg) * NT]);
        float d[8][4];
        const float* tab = cTY + 4 * cg;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            dct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float tf = tab[u * 8 + j] * f;
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
            scr[(SC_Y + 2 * r + cg) * NT] = to_f4(v[r]);
            if (GRAD) scr[(SC_DY + 2 * r + cg) * NT] = to_f4(d[r]);
        }
    }
}

// Phase 2 of the state-saving forward: as dj_luma_columns<ROUND, false, ...> but round'(q) goes
// straight to global memory (two 16-byte stores per block row; the halves of a 32-byte sector are
// written by the two column groups back to back and merge in L2).
template <int ROUND, int NT>
__device__ __forceinline__ void dj_luma_columns_save(float4* scr, float f, float* dY, int64_t W, bool active) {
#pragma unroll 1
    for (int cg = 0; cg < 2; ++cg) {
        float v[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r) f4_to(v[r], scr[(SC_Y + 2 * r + cg) * NT]);
        float d[8][4];
        const float* tab = cTY + 4 * cg;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            dct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float tf = tab[u * 8 + j] * f;
                const float q = div_by_recip(v[u][j], tf, fast_rcp(tf));
                d[u][j] = round_grad<ROUND>(q);
                v[u][j] = round_fwd<ROUND>(q) * tf;
            }
            idct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            scr[(SC_Y + 2 * r + cg) * NT] = to_f4(v[r]);
            if (active) *reinterpret_cast<float4*>(dY + int64_t(r) * W + 4 * cg) = to_f4(d[r]);
        }
    }
}

template <int ROUND, int NT>
__device__ __forceinline__ void dj_chroma_planes_save(float4* scr, const QuadCoef& qx, const QuadCoef& qy,
                                                      int bx, int by, float f, float* dC, int64_t plane_c, int64_t Wc, bool active) {
#pragma unroll 1
    for (int pl = 0; pl < 2; ++pl) {
        float p[4][4], d[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) f4_to(p[i], scr[(SC_CB + 4 * pl + i) * NT]);
        quad_dct_rows(p, qx, 1);
        quad_dct_cols(p, qy, 16);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float tf = cTC[(16 * i + 2 * j) + (8 * by + bx)] * f;   // [u=2i+by][v=2j+bx]
                const float q = div_by_recip(p[i][j], tf, fast_rcp(tf));
                d[i][j] = round_grad<ROUND>(q);
                p[i][j] = round_fwd<ROUND>(q) * tf;
            }
        quad_idct_cols(p, qy, 16);
        quad_idct_rows(p, qx, 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            scr[(SC_CB + 4 * pl + i) * NT] = to_f4(p[i]);
            if (active) *reinterpret_cast<float4*>(dC + pl * plane_c + int64_t(i) * Wc) = to_f4(d[i]);
        }
    }
}

// Phase 3: chroma quadrants (rolled over the two planes): split DCT -> quantise / round /
// dequantise -> split IDCT, in place in the SC_CB / SC_CR chunks.
template <int ROUND, bool KEEP_Q, bool GRAD, int NT>
__device__ __forceinline__ void dj_chroma_planes(float4* scr, const QuadCoef& qx, const QuadCoef& qy,
                                                 int bx, int by, float f) {
#pragma unroll 1
    for (int pl = 0; pl < 2; ++pl) {
        float p[4][4], d[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) f4_to(p[i], scr[(SC_CB + 4 * pl + i) * NT]);
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
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            scr[(SC_CB + 4 * pl + i) * NT] = to_f4(p[i]);
            if (GRAD) scr[(SC_DC + 4 * pl + i) * NT] = to_f4(d[i]);
        }
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

// torch's min/max clamp (utils/JPEG.py:467-468) propagates NaN, .sat returns 0.  A NaN (or an inf, or the 0/0 of
// quality = 100, factor 0) anywhere in an 8x8 block reaches EVERY output of the inverse transform of that block,
// so one test per row and plane is enough: rare path, nothing on the common one but two compares.
__device__ __forceinline__ void dj_propagate_nan(const float (&yv)[8], const float (&tR)[4], const float (&tB)[4],
                                                 f8& oR, f8& oG, f8& oB) {
    const float probe = yv[0] + tR[0] + tB[0] + tR[3] + tB[3];
    if (probe != probe) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float nR = yv[c] + tR[c >> 1], nB = yv[c] + tB[c >> 1], nG = nR + nB;
            if (nR != nR) oR.v[c] = nR;
            if (nG != nG) oG.v[c] = nG;
            if (nB != nB) oB.v[c] = nB;
        }
    }
}

// Final phase of forward / decompress: row IDCT, upsampled chroma, colour transform, clamp.
template <int NT, bool EP = false>
__device__ __forceinline__ void dj_emit_rgb(const DJArgs& a, const DJThread& t, const float4* scr) {
    float* yo = a.out + (int64_t(t.b) * 3 * a.H + t.row0) * a.W + t.col0;
    const int64_t plane = int64_t(a.H) * a.W;
#pragma unroll 1
    for (int rp = 0; rp < 4; ++rp) {
        float cb[4], cr[4], tR[4], tG[4], tB[4];
        f4_to(cb, scr[(SC_CB + rp) * NT]);
        f4_to(cr, scr[(SC_CR + rp) * NT]);
        dj_chroma_terms(cb, cr, tR, tG, tB);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int r = 2 * rp + rr;
            float* p = yo + int64_t(r) * a.W;
            f8 xR, xG, xB;
            if (EP && t.active) {      // store epilogue: x at the output position, requested before the row's math
                const float* xp = a.ep.x + (p - a.out);
                xR = ldg256_stream(xp); xG = ldg256_stream(xp + plane); xB = ldg256_stream(xp + 2 * plane);
            }
            float yv[8];
            scr_load_row<NT>(scr, r, yv);
            idct8(yv);
            f8 oR, oG, oB;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                // min(255, max(0, v)) / 255  (utils/JPEG.py:467-469) == saturate(v / 255)
                oR.v[c] = __saturatef(fmaf(yv[c], DJ_I255, tR[c >> 1]));
                oG.v[c] = __saturatef(fmaf(yv[c], DJ_I255, tG[c >> 1]));
                oB.v[c] = __saturatef(fmaf(yv[c], DJ_I255, tB[c >> 1]));
            }
            dj_propagate_nan(yv, tR, tB, oR, oG, oB);
            if (t.active) {
                if (EP) { ep_apply_n<8>(oR.v, xR.v, a.ep); ep_apply_n<8>(oG.v, xG.v, a.ep); ep_apply_n<8>(oB.v, xB.v, a.ep); }
                stg256(p, oR);
                stg256(p + plane, oG);
                stg256(p + 2 * plane, oB);
            }
        }
    }
}

// Final phase of the state-saving forward: as dj_emit_rgb, plus the clamp code of every value
// (0 outside, 1 strictly inside, 2 exactly on a bound: the backward multiplies by 0 / 1 / 0.5, the
// tie rule of the reference's binary min/max clamp, utils/JPEG.py:467-468): 48 bits per 8-pixel row.
__device__ __forceinline__ unsigned clamp_code(float u) {
    const bool open_in = u > 0.f && u < 1.f, closed_in = u >= 0.f && u <= 1.f;
    return open_in ? 1u : (closed_in ? 2u : 0u);
}
template <int NT>
__device__ __forceinline__ void dj_emit_rgb_save(const DJArgs& a, const DJThread& t, const float4* scr) {
    float* yo = a.out + (int64_t(t.b) * 3 * a.H + t.row0) * a.W + t.col0;
    unsigned long long* cmo = a.cm + (int64_t(t.b) * a.H + t.row0) * (a.W >> 3) + (t.col0 >> 3);
    const int64_t plane = int64_t(a.H) * a.W;
#pragma unroll 1
    for (int rp = 0; rp < 4; ++rp) {
        float cb[4], cr[4], tR[4], tG[4], tB[4];
        f4_to(cb, scr[(SC_CB + rp) * NT]);
        f4_to(cr, scr[(SC_CR + rp) * NT]);
        dj_chroma_terms(cb, cr, tR, tG, tB);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int r = 2 * rp + rr;
            float yv[8];
            scr_load_row<NT>(scr, r, yv);
            idct8(yv);
            f8 oR, oG, oB;
            // codes accumulated in base 4 as exact fp32 integers (< 2^24 per 4 pixels) with FSET + FFMA:
            // code = 2 * [s == u] - [0 < s < 1]  (1 strictly inside, 2 exactly on a bound, 0 outside)
            float accl = 0.f, acch = 0.f;     // pixels 0-3 and 4-7; pixel c occupies bits 6c .. 6c+5 (R, G, B)
#pragma unroll
            for (int c = 7; c >= 0; --c) {
                const float uR = fmaf(yv[c], DJ_I255, tR[c >> 1]), uG = fmaf(yv[c], DJ_I255, tG[c >> 1]),
                            uB = fmaf(yv[c], DJ_I255, tB[c >> 1]);
                const float sR = __saturatef(uR), sG = __saturatef(uG), sB = __saturatef(uB);
                oR.v[c] = sR; oG.v[c] = sG; oB.v[c] = sB;
                const float cB = fmaf(fset_eq(sB, uB), 2.f, -fset_gt(fmaf(-sB, sB, sB), 0.f));
                const float cG = fmaf(fset_eq(sG, uG), 2.f, -fset_gt(fmaf(-sG, sG, sG), 0.f));
                const float cR = fmaf(fset_eq(sR, uR), 2.f, -fset_gt(fmaf(-sR, sR, sR), 0.f));
                float& acc = c < 4 ? accl : acch;
                acc = fmaf(fmaf(fmaf(acc, 4.f, cB), 4.f, cG), 4.f, cR);
            }
            dj_propagate_nan(yv, tR, tB, oR, oG, oB);
            const unsigned lo = __float2uint_rn(accl), hi = __float2uint_rn(acch);
            if (t.active) {
                float* p = yo + int64_t(r) * a.W;
                stg256(p, oR);
                stg256(p + plane, oG);
                stg256(p + 2 * plane, oB);
                cmo[int64_t(r) * (a.W >> 3)] = (unsigned long long)lo | ((unsigned long long)hi << 32);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
static inline int dj_check(const void* x, int64_t sb, int64_t sc, int64_t sh, int B, int H, int W, const char* who,
                           int dt = WM_DT_F32) {
    WM_REQUIRE(x != nullptr, WM_E_NULL, "%s: null image pointer", who);
    WM_REQUIRE(dtype_ok(dt), WM_E_ARG, "%s: unknown element type %d (WM_DT_F32 / WM_DT_F16 / WM_DT_BF16)", who, dt);
    WM_REQUIRE(B >= 0 && H > 0 && W > 0 && H % 16 == 0 && W % 16 == 0, WM_E_SHAPE,
               "%s: H and W must be positive multiples of 16 (got B=%d H=%d W=%d); the reference's "
               "block_merging views require it (utils/JPEG.py:371-376)", who, B, H, W);
    WM_REQUIRE(aligned(x, 8 * dtype_size(dt)) && sb % 8 == 0 && sc % 8 == 0 && sh % 8 == 0, WM_E_ALIGN,
               "%s: base pointer must be aligned to 8 elements (32 bytes for float32) and strides multiples of 8 elements "
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
static inline int dj_launch(K kernel, const DJArgs& a, int threads, size_t smem, cudaStream_t st, const char* who) {
    if (a.n_mcu == 0) return WM_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, who);
    const int64_t warps = (a.n_mcu + 7) / 8;
    const int64_t blocks = (warps * 32 + threads - 1) / threads;
    kernel<<<(unsigned)blocks, threads, smem, st>>>(a);
    WM_LAUNCH_CHECK(who);
    return WM_OK;
}

}  // namespace wm

#define DJ_DISPATCH_ROUND_T(KERNEL, TYPED, ...)                                               \
    switch (rounding) {                                                                       \
        case WM_ROUND_ONLY_AT_0: return dj_launch(KERNEL<WM_ROUND_ONLY_AT_0, TYPED>, __VA_ARGS__);   \
        case WM_ROUND_CUBIC:     return dj_launch(KERNEL<WM_ROUND_CUBIC, TYPED>, __VA_ARGS__);       \
        case WM_ROUND_HARD:      return dj_launch(KERNEL<WM_ROUND_HARD, TYPED>, __VA_ARGS__);        \
        case WM_ROUND_FOURIER:   return dj_launch(KERNEL<WM_ROUND_FOURIER, TYPED>, __VA_ARGS__);     \
        default: set_error("unknown rounding mode %d", rounding); return WM_E_ARG;            \
    }

#define DJ_DISPATCH_ROUND(KERNEL, ...)                                                        \
    switch (rounding) {                                                                       \
        case WM_ROUND_ONLY_AT_0: return dj_launch(KERNEL<WM_ROUND_ONLY_AT_0>, __VA_ARGS__);   \
        case WM_ROUND_CUBIC:     return dj_launch(KERNEL<WM_ROUND_CUBIC>, __VA_ARGS__);       \
        case WM_ROUND_HARD:      return dj_launch(KERNEL<WM_ROUND_HARD>, __VA_ARGS__);        \
        case WM_ROUND_FOURIER:   return dj_launch(KERNEL<WM_ROUND_FOURIER>, __VA_ARGS__);     \
        default: set_error("unknown rounding mode %d", rounding); return WM_E_ARG;            \
    }

