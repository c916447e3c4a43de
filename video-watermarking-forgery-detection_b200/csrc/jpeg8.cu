// 8x8-unit 4:4:4 JPEG simulators: Jpeg / JpegSS / JpegMask (noise_layers/jpeg.py:214-306) and
// HiDDeN JpegCompression (noise_layers/jpeg_compression.py:65-159) in one templated kernel.
//
// Reference cost replaced: 155-174 materialising torch kernels per forward+backward (table
// rebuilds, split/cat re-blocking, 56 scalar device writes to build `coff`, per-channel conv2d).
// Here: one thread owns one 8x8 pixel block (3 channels); Y and U stream through a
// thread-private shared-memory scratch exactly like the DiffJPEG luminance block, V stays in
// registers.  Zero padding to a multiple of 8 (jpeg.py:171-173) is done by predicated loads,
// un-padding by predicated stores.  HiDDeN's un-normalised conv DCT followed by its matching
// synthesis filters equals the orthonormal DCT -> 0/1 mask -> inverse, so it runs as MASK.
#include "dct8.cuh"
#include "wm_common.cuh"

namespace wm {

constexpr int J8_THREADS = 128;
constexpr int J8B_THREADS = 64;

struct J8Args {
    const float* x; int64_t x_sb, x_sc, x_sh;
    const float* gy; int64_t g_sb, g_sc, g_sh;
    float* out;        // y / gx, dense [B,3,H,W]
    int x_dt, out_dt;  // element types of x and out (WM_DT_*): the two-threads-per-block kernel reads / stores them directly
    float* coef;       // optional [B,3,Hp,Wp] quantised coefficient image
    int B, H, W, Hb, Wb; int64_t n_blk;
    float fwd[9], inv[9];
    float table[3][64];
    float rtable[3][64];
    StoreEp ep;
};

struct J8Thread { bool active; int b, row0, col0; };

__device__ __forceinline__ J8Thread j8_locate(const J8Args& a) {
    J8Thread t;
    const int64_t blk = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    t.active = blk < a.n_blk;
    const int64_t m = t.active ? blk : 0;
    const int per = a.Hb * a.Wb;
    t.b = int(m / per);
    const int rem = int(m - int64_t(t.b) * per);
    const int by = rem / a.Wb;
    t.row0 = by * 8; t.col0 = (rem - by * a.Wb) * 8;
    return t;
}

// predicated row load: 8 floats at p (valid count `n` <= 8), zero elsewhere
template <bool VEC>
__device__ __forceinline__ void j8_load_row(const float* p, bool row_ok, int n, float (&v)[8]) {
    if (VEC) {
        if (row_ok) { f8 t = ldg256_stream(p);
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = t.v[c]; }
        else {
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = 0.f; }
    } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = (row_ok && c < n) ? __ldg(p + c) : 0.f;
    }
}
template <bool VEC>
__device__ __forceinline__ void j8_store_row(float* p, bool row_ok, int n, const float (&v)[8]) {
    if (VEC) {
        if (row_ok) { f8 t;
#pragma unroll
            for (int c = 0; c < 8; ++c) t.v[c] = v[c];
            stg256(p, t); }
    } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) if (row_ok && c < n) p[c] = v[c];
    }
}

template <int THREADS>
__device__ __forceinline__ void j8_scr_store(float4* scr, int r, const float (&v)[8]) {
    scr[(2 * r) * THREADS] = make_float4(v[0], v[1], v[2], v[3]);
    scr[(2 * r + 1) * THREADS] = make_float4(v[4], v[5], v[6], v[7]);
}
template <int THREADS>
__device__ __forceinline__ void j8_scr_load(const float4* scr, int r, float (&v)[8]) {
    float4 a = scr[(2 * r) * THREADS], b = scr[(2 * r + 1) * THREADS];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// JpegSS.round_ss (noise_layers/jpeg.py:255-257) == round_only_at_0
__device__ __forceinline__ float ss_fwd(float q) { return fabsf(q) < 0.5f ? q * q * q : q; }
__device__ __forceinline__ float ss_grad(float q) { return fabsf(q) < 0.5f ? 3.f * q * q : 1.f; }

// quantise / round / dequantise one coefficient of channel CH at (u, v)
//   OUT 0: dequantised value   OUT 1: rounded quantised value   OUT 2: d round / dq
template <int VARIANT, int OUT>
__device__ __forceinline__ float j8_quant(const J8Args& a, int ch, int uv, float c) {
    const float T = a.table[ch][uv];
    if (VARIANT == WM_JPEG8_MASK) return OUT == 2 ? T : c * T;
    const float q = div_by_recip(c, T, a.rtable[ch][uv]);
    if (OUT == 2) return VARIANT == WM_JPEG8_SS ? ss_grad(q) : 0.f;
    const float r = VARIANT == WM_JPEG8_SS ? ss_fwd(q) : rintf(q);
    return OUT == 1 ? r : r * T;
}

// column stage of one channel held in scratch (4-column groups)
template <int VARIANT, int OUT, int THREADS>
__device__ __forceinline__ void j8_columns_scr(const J8Args& a, int ch, float4* scr) {
#pragma unroll
    for (int cg = 0; cg < 2; ++cg) {
        float v[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float4 t4 = scr[(2 * r + cg) * THREADS];
            v[r][0] = t4.x; v[r][1] = t4.y; v[r][2] = t4.z; v[r][3] = t4.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            dct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u][j] = j8_quant<VARIANT, OUT>(a, ch, u * 8 + 4 * cg + j, v[u][j]);
            if (OUT == 0) idct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
            scr[(2 * r + cg) * THREADS] = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
    }
}

// SUBMODE 0: none; 2: replicate even rows/cols onto odd ones before the DCT (jpeg.py:202-211);
//         3: the adjoint of 2 applied after the inverse DCT (fold odd onto even, zero odd).
template <int VARIANT, int SUBMODE, bool VEC, bool QOUT>
__global__ void __launch_bounds__(J8_THREADS, 3) jpeg8_fwd_kernel(const J8Args a) {
    extern __shared__ float4 smem[];
    float4* sY = smem + threadIdx.x;
    float4* sU = smem + 16 * J8_THREADS + threadIdx.x;
    const J8Thread t = j8_locate(a);
    const float* xr = a.x + int64_t(t.b) * a.x_sb + int64_t(t.row0) * a.x_sh + t.col0;
    const int ncol = min(8, a.W - t.col0);
    float vv[8][8];                       // third channel, register resident
    float uprev[8], vprev[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const bool ok = t.active && (t.row0 + r) < a.H;
        const float* p = xr + int64_t(r) * a.x_sh;
        float R[8], G[8], Bl[8];
        j8_load_row<VEC>(p, ok, ncol, R);
        j8_load_row<VEC>(p + a.x_sc, ok, ncol, G);
        j8_load_row<VEC>(p + 2 * a.x_sc, ok, ncol, Bl);
        float y[8], u[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            y[c] = fmaf(a.fwd[0], R[c], fmaf(a.fwd[1], G[c], a.fwd[2] * Bl[c]));
            u[c] = fmaf(a.fwd[3], R[c], fmaf(a.fwd[4], G[c], a.fwd[5] * Bl[c]));
            vv[r][c] = fmaf(a.fwd[6], R[c], fmaf(a.fwd[7], G[c], a.fwd[8] * Bl[c]));
        }
        if (SUBMODE == 2) {
            if (r & 1) {
#pragma unroll
                for (int c = 0; c < 8; ++c) { u[c] = uprev[c]; vv[r][c] = vprev[c]; }
            } else {
#pragma unroll
                for (int c = 1; c < 8; c += 2) { u[c] = u[c - 1]; vv[r][c] = vv[r][c - 1]; }
#pragma unroll
                for (int c = 0; c < 8; ++c) { uprev[c] = u[c]; vprev[c] = vv[r][c]; }
            }
        }
        dct8(y); dct8(u); dct8(vv[r]);
        j8_scr_store<J8_THREADS>(sY, r, y);
        j8_scr_store<J8_THREADS>(sU, r, u);
    }
    constexpr int OUT = QOUT ? 1 : 0;
    j8_columns_scr<VARIANT, OUT, J8_THREADS>(a, 0, sY);
    j8_columns_scr<VARIANT, OUT, J8_THREADS>(a, 1, sU);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        dct8(vv[0][c], vv[1][c], vv[2][c], vv[3][c], vv[4][c], vv[5][c], vv[6][c], vv[7][c]);
#pragma unroll
        for (int u = 0; u < 8; ++u) vv[u][c] = j8_quant<VARIANT, OUT>(a, 2, u * 8 + c, vv[u][c]);
        if (!QOUT) idct8(vv[0][c], vv[1][c], vv[2][c], vv[3][c], vv[4][c], vv[5][c], vv[6][c], vv[7][c]);
    }
    if (QOUT) {
        // coefficient image [B,3,Hp,Wp], block (i,j) coefficient (u,v) at [8i+u, 8j+v]
        const int Hp = a.Hb * 8, Wp = a.Wb * 8;
        float* co = a.coef + (int64_t(t.b) * 3 * Hp + t.row0) * Wp + t.col0;
        const int64_t plane = int64_t(Hp) * Wp;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float y[8], u[8];
            j8_scr_load<J8_THREADS>(sY, r, y);
            j8_scr_load<J8_THREADS>(sU, r, u);
            j8_store_row<true>(co + int64_t(r) * Wp, t.active, 8, y);
            j8_store_row<true>(co + plane + int64_t(r) * Wp, t.active, 8, u);
            j8_store_row<true>(co + 2 * plane + int64_t(r) * Wp, t.active, 8, vv[r]);
        }
        return;
    }
    float* yo = a.out + (int64_t(t.b) * 3 * a.H + t.row0) * a.W + t.col0;
    const int64_t plane = int64_t(a.H) * a.W;
    float ufold[8], vfold[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        float y[8], u[8];
        j8_scr_load<J8_THREADS>(sY, r, y);
        j8_scr_load<J8_THREADS>(sU, r, u);
        idct8(y); idct8(u); idct8(vv[r]);
        if (SUBMODE == 3) {
            // adjoint of the replicate: even (row, col) collects its 2x2 cell, the rest is zero.
            // Rows are visited in order, so fold the odd row of the NEXT iteration early:
            // process pairs by looking ahead (r even: stash; r odd: emit both rows).
            if ((r & 1) == 0) {
#pragma unroll
                for (int c = 0; c < 8; ++c) { ufold[c] = u[c]; vfold[c] = vv[r][c]; }
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) { ufold[c] += u[c]; vfold[c] += vv[r][c]; u[c] = 0.f; vv[r][c] = 0.f; }
#pragma unroll
                for (int c = 0; c < 8; c += 2) { ufold[c] += ufold[c + 1]; vfold[c] += vfold[c + 1]; ufold[c + 1] = 0.f; vfold[c + 1] = 0.f; }
            }
        }
        if (SUBMODE == 3 && (r & 1) == 0) {
            // even row: its chroma contribution is only known after the odd row; park Y
            j8_scr_store<J8_THREADS>(sY, r, y);
            continue;
        }
        const int nrows = (SUBMODE == 3) ? 2 : 1;
#pragma unroll
        for (int k = 0; k < nrows; ++k) {
            const int rr = (SUBMODE == 3) ? r - 1 + k : r;
            float yy[8], uu[8], ww[8];
            if (SUBMODE == 3 && k == 0) {
                j8_scr_load<J8_THREADS>(sY, rr, yy);
#pragma unroll
                for (int c = 0; c < 8; ++c) { uu[c] = ufold[c]; ww[c] = vfold[c]; }
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) { yy[c] = y[c]; uu[c] = u[c]; ww[c] = vv[r][c]; }
            }
            float oR[8], oG[8], oB[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                oR[c] = fmaf(a.inv[0], yy[c], fmaf(a.inv[1], uu[c], a.inv[2] * ww[c]));
                oG[c] = fmaf(a.inv[3], yy[c], fmaf(a.inv[4], uu[c], a.inv[5] * ww[c]));
                oB[c] = fmaf(a.inv[6], yy[c], fmaf(a.inv[7], uu[c], a.inv[8] * ww[c]));
            }
            const bool ok = t.active && (t.row0 + rr) < a.H;
            float* p = yo + int64_t(rr) * a.W;
            j8_store_row<VEC>(p, ok, ncol, oR);
            j8_store_row<VEC>(p + plane, ok, ncol, oG);
            j8_store_row<VEC>(p + 2 * plane, ok, ncol, oB);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fast forward path (aligned rows, W % 8 == 0, no in-block chroma decimation): TWO threads per
// 8x8 block.  Lanes 0-15 of a warp own rows 0-3 / the left 4 columns of 16 consecutive blocks,
// lanes 16-31 rows 4-7 / the right 4 columns of the same blocks; all three channels of a block
// stream through 48 shared-memory chunks, so a thread never holds more than one row triple or
// one 8x4 column group: ~half the registers and code of the one-thread-per-block kernel, twice
// the warps per SM (16), and the column stage is a rolled loop over the channels.  Quarter-warps
// touch 8 different blocks with the same chunk index: conflict-free LDS.128 / STS.128.
// ---------------------------------------------------------------------------------------------
constexpr int J8P_THREADS = 128, J8P_BLOCKS = 64, J8P_CHUNKS = 48;
#ifndef WM_J8_LOAD_UNROLL
#define WM_J8_LOAD_UNROLL 2
#endif
constexpr int J8P_LOAD_UNROLL = WM_J8_LOAD_UNROLL;      // rows of the load + row-DCT phase in flight per thread

// DMODE 0: plain forward.  1: JpegSS forward that also saves ss'(q) of every coefficient (12 B/px,
// [B,3,Hp,W] in coefficient-image order).  2: JpegSS backward from that state: the cotangent runs
// through inv_color^T -> DCT -> times ss'(q) -> IDCT -> fwd_color^T, i.e. the linear pipeline with a
// per-coefficient "mask" read from global memory (the quantisation steps cancel); no recompute.
// VEC = false: rows that are not 32-byte aligned or a width that is not a multiple of 8 (scalar, predicated row
// access; the missing columns / rows of the last blocks are the zero padding of noise_layers/jpeg.py:171-173).
// TYPED = false: float32 in and out, the element-type switches fold away (the kernel of the headline path).
template <int VARIANT, int DMODE, bool VEC = true, bool TYPED = false>
__global__ void __launch_bounds__(J8P_THREADS, 4) jpeg8_pair_kernel(const J8Args a) {
    const int xdt = TYPED ? a.x_dt : WM_DT_F32, odt = TYPED ? a.out_dt : WM_DT_F32;
    extern __shared__ float4 smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = lane >> 4, bic = warp * 16 + (lane & 15);          // half, block in CTA
    float4* scr = smem + bic;                                         // chunk c at scr[c * J8P_BLOCKS]
    const int64_t blk = int64_t(blockIdx.x) * J8P_BLOCKS + bic;
    const bool active = blk < a.n_blk;
    const int64_t m = active ? blk : 0;
    const int per = a.Hb * a.Wb;
    const int b = int(m / per), rem = int(m - int64_t(b) * per), by = rem / a.Wb;
    const int row0 = by * 8, col0 = (rem - by * a.Wb) * 8;
    const int64_t xo = int64_t(b) * a.x_sb + int64_t(row0 + 4 * h) * a.x_sh + col0;       // element offset (a.x_dt elements)
    const float* xr = a.x + xo;
    const int ncol = min(8, a.W - col0);

    // ---- rows 4h .. 4h+3: colour transform + row DCT of the three channels -> scratch --------------
#pragma unroll J8P_LOAD_UNROLL
    for (int i = 0; i < 4; ++i) {
        const int r = 4 * h + i;
        const bool ok = active && (row0 + r) < a.H;
        const float* p = xr + int64_t(i) * a.x_sh;
        float R[8], G[8], Bl[8];
        if (VEC) {          // float32 / float16 / bfloat16 rows, 8 elements per load
            const int64_t po = xo + int64_t(i) * a.x_sh;
            f8 tr{}, tg{}, tb{};
            if (ok) { tr = ld8_typed(a.x, po, xdt); tg = ld8_typed(a.x, po + a.x_sc, xdt); tb = ld8_typed(a.x, po + 2 * a.x_sc, xdt); }
#pragma unroll
            for (int c = 0; c < 8; ++c) { R[c] = tr.v[c]; G[c] = tg.v[c]; Bl[c] = tb.v[c]; }
        } else {
            j8_load_row<false>(p, ok, ncol, R);
            j8_load_row<false>(p + a.x_sc, ok, ncol, G);
            j8_load_row<false>(p + 2 * a.x_sc, ok, ncol, Bl);
        }
        float y[8], u[8], v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            y[c] = fmaf(a.fwd[0], R[c], fmaf(a.fwd[1], G[c], a.fwd[2] * Bl[c]));
            u[c] = fmaf(a.fwd[3], R[c], fmaf(a.fwd[4], G[c], a.fwd[5] * Bl[c]));
            v[c] = fmaf(a.fwd[6], R[c], fmaf(a.fwd[7], G[c], a.fwd[8] * Bl[c]));
        }
        dct8(y); dct8(u); dct8(v);
        j8_scr_store<J8P_BLOCKS>(scr, r, y);
        j8_scr_store<J8P_BLOCKS>(scr + 16 * J8P_BLOCKS, r, u);
        j8_scr_store<J8P_BLOCKS>(scr + 32 * J8P_BLOCKS, r, v);
    }
    __syncwarp();

    // ---- columns 4h .. 4h+3 of every channel: DCT, quantise / round / dequantise, IDCT --------------
#pragma unroll 1
    for (int ch = 0; ch < 3; ++ch) {
        float4* sc = scr + (16 * ch) * J8P_BLOCKS;
        float v[8][4];
        float d[8][4];
        // saved state: coefficient (u, v) of this block at [b, ch, row0 + u, col0 + v]
        float* dp = DMODE ? a.coef + ((int64_t(b) * 3 + ch) * (a.Hb * 8) + row0) * a.W + col0 + 4 * h : nullptr;
        if (DMODE == 2) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float4 t4 = active ? ldg128_stream(dp + int64_t(r) * a.W) : make_float4(0.f, 0.f, 0.f, 0.f);
                d[r][0] = t4.x; d[r][1] = t4.y; d[r][2] = t4.z; d[r][3] = t4.w;
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float4 t4 = sc[(2 * r + h) * J8P_BLOCKS];
            v[r][0] = t4.x; v[r][1] = t4.y; v[r][2] = t4.z; v[r][3] = t4.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            dct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (DMODE == 2) { v[u][j] *= d[u][j]; continue; }
                if (DMODE == 1) d[u][j] = j8_quant<VARIANT, 2>(a, ch, u * 8 + 4 * h + j, v[u][j]);
                v[u][j] = j8_quant<VARIANT, 0>(a, ch, u * 8 + 4 * h + j, v[u][j]);
            }
            idct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
        }
        if (DMODE == 1 && active) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
                *reinterpret_cast<float4*>(dp + int64_t(r) * a.W) = make_float4(d[r][0], d[r][1], d[r][2], d[r][3]);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) sc[(2 * r + h) * J8P_BLOCKS] = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
    }
    __syncwarp();

    // ---- rows 4h .. 4h+3: row IDCT, inverse colour transform, store -----------------------------
    float* yo = a.out + (int64_t(b) * 3 * a.H + row0 + 4 * h) * a.W + col0;
    const int64_t plane = int64_t(a.H) * a.W;
#pragma unroll 2
    for (int i = 0; i < 4; ++i) {
        const int r = 4 * h + i;
        const bool ok = active && (row0 + r) < a.H;
        float* p = yo + int64_t(i) * a.W;
        f8 xR, xG, xB;
        if (a.ep.x && ok) {          // store epilogue: x at the output position, requested before the row's math
            const float* xp = a.ep.x + (p - a.out);
            xR = ldg256_stream(xp); xG = ldg256_stream(xp + plane); xB = ldg256_stream(xp + 2 * plane);
        }
        float y[8], u[8], v[8];
        j8_scr_load<J8P_BLOCKS>(scr, r, y);
        j8_scr_load<J8P_BLOCKS>(scr + 16 * J8P_BLOCKS, r, u);
        j8_scr_load<J8P_BLOCKS>(scr + 32 * J8P_BLOCKS, r, v);
        idct8(y); idct8(u); idct8(v);
        float oR[8], oG[8], oB[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            oR[c] = fmaf(a.inv[0], y[c], fmaf(a.inv[1], u[c], a.inv[2] * v[c]));
            oG[c] = fmaf(a.inv[3], y[c], fmaf(a.inv[4], u[c], a.inv[5] * v[c]));
            oB[c] = fmaf(a.inv[6], y[c], fmaf(a.inv[7], u[c], a.inv[8] * v[c]));
        }
        if (a.ep.x && ok) { ep_apply_n<8>(oR, xR.v, a.ep); ep_apply_n<8>(oG, xG.v, a.ep); ep_apply_n<8>(oB, xB.v, a.ep); }   // (dense, same layout as out)
        if (VEC) {
            if (ok) {
                const int64_t po = p - a.out;                    // element offset, a.out_dt elements
                f8 tr, tg, tb;
#pragma unroll
                for (int c = 0; c < 8; ++c) { tr.v[c] = oR[c]; tg.v[c] = oG[c]; tb.v[c] = oB[c]; }
                st8_typed(a.out, po, tr, odt); st8_typed(a.out, po + plane, tg, odt); st8_typed(a.out, po + 2 * plane, tb, odt);
            }
        } else {
            j8_store_row<false>(p, ok, ncol, oR);
            j8_store_row<false>(p + plane, ok, ncol, oG);
            j8_store_row<false>(p + 2 * plane, ok, ncol, oB);
        }
    }
}

template <int VARIANT, int DMODE, bool VEC, bool TYPED>
static int j8_pair_launch_t(const J8Args& a, cudaStream_t st, const char* who) {
    if (a.n_blk == 0) return WM_OK;
    const size_t smem = size_t(J8P_CHUNKS) * J8P_BLOCKS * sizeof(float4);
    cudaError_t e = cudaFuncSetAttribute(jpeg8_pair_kernel<VARIANT, DMODE, VEC, TYPED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, who);
    const int64_t blocks = (a.n_blk + J8P_BLOCKS - 1) / J8P_BLOCKS;
    jpeg8_pair_kernel<VARIANT, DMODE, VEC, TYPED><<<(unsigned)blocks, J8P_THREADS, smem, st>>>(a);
    WM_LAUNCH_CHECK(who);
    return WM_OK;
}
template <int VARIANT, int DMODE = 0, bool VEC = true>
static int j8_pair_launch(const J8Args& a, cudaStream_t st, const char* who) {
    if (VEC && (a.x_dt != WM_DT_F32 || a.out_dt != WM_DT_F32)) return j8_pair_launch_t<VARIANT, DMODE, VEC, VEC>(a, st, who);
    return j8_pair_launch_t<VARIANT, DMODE, VEC, false>(a, st, who);
}

// ---------------------------------------------------------------------------------------------
// JpegSS backward: per channel, recompute round'(q) from x, push the cotangent through
// DCT -> *round' -> IDCT (the quantisation steps cancel), park, then apply fwd_color^T.
// ---------------------------------------------------------------------------------------------
template <int SUBMODE, bool VEC>
__global__ void __launch_bounds__(J8B_THREADS, 2) jpeg8_ss_bwd_kernel(const J8Args a) {
    extern __shared__ float4 smem[];
    float4* sD = smem + threadIdx.x;                              // round'(q) of the channel
    float4* sG = smem + 16 * J8B_THREADS + threadIdx.x;           // cotangent working block
    float4* sP = smem + 32 * J8B_THREADS + threadIdx.x;           // 3 parked channel gradients
    const J8Thread t = j8_locate(a);
    const float* xr = a.x + int64_t(t.b) * a.x_sb + int64_t(t.row0) * a.x_sh + t.col0;
    const float* gr = a.gy + int64_t(t.b) * a.g_sb + int64_t(t.row0) * a.g_sh + t.col0;
    const int ncol = min(8, a.W - t.col0);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float prev[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const bool ok = t.active && (t.row0 + r) < a.H;
            const float* p = xr + int64_t(r) * a.x_sh;
            float R[8], G[8], Bl[8], v[8];
            j8_load_row<VEC>(p, ok, ncol, R);
            j8_load_row<VEC>(p + a.x_sc, ok, ncol, G);
            j8_load_row<VEC>(p + 2 * a.x_sc, ok, ncol, Bl);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                v[c] = fmaf(a.fwd[3 * ch], R[c], fmaf(a.fwd[3 * ch + 1], G[c], a.fwd[3 * ch + 2] * Bl[c]));
            if (SUBMODE == 2 && ch > 0) {
                if (r & 1) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[c] = prev[c];
                } else {
#pragma unroll
                    for (int c = 1; c < 8; c += 2) v[c] = v[c - 1];
#pragma unroll
                    for (int c = 0; c < 8; ++c) prev[c] = v[c];
                }
            }
            dct8(v);
            j8_scr_store<J8B_THREADS>(sD, r, v);
        }
        j8_columns_scr<WM_JPEG8_SS, 2, J8B_THREADS>(a, ch, sD);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const bool ok = t.active && (t.row0 + r) < a.H;
            const float* p = gr + int64_t(r) * a.g_sh;
            float R[8], G[8], Bl[8], v[8];
            j8_load_row<VEC>(p, ok, ncol, R);
            j8_load_row<VEC>(p + a.g_sc, ok, ncol, G);
            j8_load_row<VEC>(p + 2 * a.g_sc, ok, ncol, Bl);
#pragma unroll
            for (int c = 0; c < 8; ++c)      // inv_color^T, row `ch`
                v[c] = fmaf(a.inv[ch], R[c], fmaf(a.inv[3 + ch], G[c], a.inv[6 + ch] * Bl[c]));
            dct8(v);
            j8_scr_store<J8B_THREADS>(sG, r, v);
        }
#pragma unroll
        for (int cg = 0; cg < 2; ++cg) {
            float v[8][4];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float4 t4 = sG[(2 * r + cg) * J8B_THREADS];
                v[r][0] = t4.x; v[r][1] = t4.y; v[r][2] = t4.z; v[r][3] = t4.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                dct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float4 d4 = sD[(2 * r + cg) * J8B_THREADS];
                v[r][0] *= d4.x; v[r][1] *= d4.y; v[r][2] *= d4.z; v[r][3] *= d4.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                idct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
#pragma unroll
            for (int r = 0; r < 8; ++r)
                sG[(2 * r + cg) * J8B_THREADS] = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
        }
        float4* park = sP + ch * 16 * J8B_THREADS;
        float fold[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float v[8];
            j8_scr_load<J8B_THREADS>(sG, r, v);
            idct8(v);
            if (SUBMODE == 2 && ch > 0) {
                if ((r & 1) == 0) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) fold[c] = v[c];
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) { fold[c] += v[c]; v[c] = 0.f; }
#pragma unroll
                    for (int c = 0; c < 8; c += 2) { fold[c] += fold[c + 1]; fold[c + 1] = 0.f; }
                    j8_scr_store<J8B_THREADS>(park, r - 1, fold);
                    j8_scr_store<J8B_THREADS>(park, r, v);
                }
            } else {
                j8_scr_store<J8B_THREADS>(park, r, v);
            }
        }
    }
    float* go = a.out + (int64_t(t.b) * 3 * a.H + t.row0) * a.W + t.col0;
    const int64_t plane = int64_t(a.H) * a.W;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        float y[8], u[8], v[8], oR[8], oG[8], oB[8];
        j8_scr_load<J8B_THREADS>(sP, r, y);
        j8_scr_load<J8B_THREADS>(sP + 16 * J8B_THREADS, r, u);
        j8_scr_load<J8B_THREADS>(sP + 32 * J8B_THREADS, r, v);
#pragma unroll
        for (int c = 0; c < 8; ++c) {      // fwd_color^T
            oR[c] = fmaf(a.fwd[0], y[c], fmaf(a.fwd[3], u[c], a.fwd[6] * v[c]));
            oG[c] = fmaf(a.fwd[1], y[c], fmaf(a.fwd[4], u[c], a.fwd[7] * v[c]));
            oB[c] = fmaf(a.fwd[2], y[c], fmaf(a.fwd[5], u[c], a.fwd[8] * v[c]));
        }
        const bool ok = t.active && (t.row0 + r) < a.H;
        float* p = go + int64_t(r) * a.W;
        j8_store_row<VEC>(p, ok, ncol, oR);
        j8_store_row<VEC>(p + plane, ok, ncol, oG);
        j8_store_row<VEC>(p + 2 * plane, ok, ncol, oB);
    }
}

// ---------------------------------------------------------------------------------------------
static int j8_fill(J8Args& a, const float* x, int64_t sb, int64_t sc, int64_t sh, int B, int H, int W,
                   const wm_jpeg8_params* p, const char* who) {
    WM_REQUIRE(x != nullptr && p != nullptr, WM_E_NULL, "%s: null pointer", who);
    WM_REQUIRE(B >= 0 && H > 0 && W > 0, WM_E_SHAPE, "%s: bad shape B=%d H=%d W=%d", who, B, H, W);
    WM_REQUIRE(p->variant >= 0 && p->variant <= 2 && (p->subsample == 0 || p->subsample == 2), WM_E_ARG,
               "%s: variant must be 0..2 and subsample 0 or 2 (got %d, %d)", who, p->variant, p->subsample);
    a.x = x; a.x_sb = sb; a.x_sc = sc; a.x_sh = sh;
    a.B = B; a.H = H; a.W = W; a.Hb = (H + 7) / 8; a.Wb = (W + 7) / 8;
    a.n_blk = int64_t(B) * a.Hb * a.Wb;
    for (int i = 0; i < 9; ++i) { a.fwd[i] = p->fwd_color[i]; a.inv[i] = p->inv_color[i]; }
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < 64; ++i) {
            const float T = p->table[c][i];
            WM_REQUIRE(p->variant == WM_JPEG8_MASK || T > 0.f, WM_E_ARG, "%s: quantisation step must be > 0", who);
            a.table[c][i] = T;
            a.rtable[c][i] = T != 0.f ? (float)(1.0 / (double)T) : 0.f;
        }
    return WM_OK;
}

static bool j8_vec_ok(const void* p, int64_t sb, int64_t sc, int64_t sh, int W, int dt = WM_DT_F32) {
    return aligned(p, 8 * dtype_size(dt)) && W % 8 == 0 && sb % 8 == 0 && sc % 8 == 0 && sh % 8 == 0;
}

template <typename K>
static int j8_launch(K kernel, const J8Args& a, int threads, size_t smem, cudaStream_t st, const char* who) {
    if (a.n_blk == 0) return WM_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, who);
    const int64_t blocks = (a.n_blk + threads - 1) / threads;
    kernel<<<(unsigned)blocks, threads, smem, st>>>(a);
    WM_LAUNCH_CHECK(who);
    return WM_OK;
}

template <int VARIANT, int SUBMODE>
static int j8_fwd_dispatch(const J8Args& a, bool vec, bool qout, cudaStream_t st, const char* who) {
    const size_t smem = 32 * J8_THREADS * sizeof(float4);
    if (qout) return vec ? j8_launch(jpeg8_fwd_kernel<VARIANT, SUBMODE, true, true>, a, J8_THREADS, smem, st, who)
                         : j8_launch(jpeg8_fwd_kernel<VARIANT, SUBMODE, false, true>, a, J8_THREADS, smem, st, who);
    return vec ? j8_launch(jpeg8_fwd_kernel<VARIANT, SUBMODE, true, false>, a, J8_THREADS, smem, st, who)
               : j8_launch(jpeg8_fwd_kernel<VARIANT, SUBMODE, false, false>, a, J8_THREADS, smem, st, who);
}

static int j8_fwd_any(const J8Args& a, int variant, int submode, bool vec, bool qout, cudaStream_t st, const char* who) {
    if (!qout && submode == 0) {                 // two threads per block; ragged rows take its scalar-access instantiation
        if (variant == WM_JPEG8_HARD) return vec ? j8_pair_launch<WM_JPEG8_HARD>(a, st, who) : j8_pair_launch<WM_JPEG8_HARD, 0, false>(a, st, who);
        if (variant == WM_JPEG8_SS) return vec ? j8_pair_launch<WM_JPEG8_SS>(a, st, who) : j8_pair_launch<WM_JPEG8_SS, 0, false>(a, st, who);
        if (variant == WM_JPEG8_MASK) return vec ? j8_pair_launch<WM_JPEG8_MASK>(a, st, who) : j8_pair_launch<WM_JPEG8_MASK, 0, false>(a, st, who);
    }
#define J8_CASE(V, S) if (variant == V && submode == S) return j8_fwd_dispatch<V, S>(a, vec, qout, st, who);
    J8_CASE(WM_JPEG8_HARD, 0) J8_CASE(WM_JPEG8_HARD, 2)
    J8_CASE(WM_JPEG8_SS, 0) J8_CASE(WM_JPEG8_SS, 2)
    J8_CASE(WM_JPEG8_MASK, 0) J8_CASE(WM_JPEG8_MASK, 2) J8_CASE(WM_JPEG8_MASK, 3)
#undef J8_CASE
    set_error("%s: unsupported variant/subsample combination (%d, %d)", who, variant, submode);
    return WM_E_ARG;
}

}  // namespace wm

using namespace wm;

// typed x / gx (WM_DT_F16 / WM_DT_BF16) exist on the vector path of the two-threads-per-block kernel only
#define J8_TYPED_CHECK(dt, ok, who)                                                                                    \
    WM_REQUIRE(dtype_ok(dt), WM_E_ARG, "%s: unknown element type %d", who, dt);                                         \
    WM_REQUIRE((dt) == WM_DT_F32 || (ok), WM_E_ARG,                                                                     \
               "%s: float16 / bfloat16 tensors need the vector path (W %% 8 == 0, no subsampling, pointers aligned to 8 elements)", who)

extern "C" int wm_jpeg8_fwd(const void* x, int x_dtype, int64_t x_sb, int64_t x_sc, int64_t x_sh, float* y,
                            int B, int H, int W, const wm_jpeg8_params* p, const wm_store_epilogue* ep, void* stream) {
    if (B == 0) return WM_OK;      // empty work: nothing to validate or launch
    J8Args a{};
    if (int rc = j8_fill(a, reinterpret_cast<const float*>(x), x_sb, x_sc, x_sh, B, H, W, p, "wm_jpeg8_fwd")) return rc;
    WM_EP_CHECK(ep, "wm_jpeg8_fwd");
    WM_REQUIRE(y != nullptr, WM_E_NULL, "wm_jpeg8_fwd: null output");
    a.out = y;
    const bool vec = dtype_ok(x_dtype) && j8_vec_ok(x, x_sb, x_sc, x_sh, W, x_dtype) && aligned(y, 32);
    J8_TYPED_CHECK(x_dtype, vec && p->subsample == 0, "wm_jpeg8_fwd");
    a.x_dt = x_dtype;
    if (vec && p->subsample == 0) a.ep = make_store_ep(ep);           // the two-threads-per-block kernel applies it
    else WM_EP_REJECT(ep, "wm_jpeg8_fwd (ragged / subsampled path)");
    return j8_fwd_any(a, p->variant, p->subsample, vec, false, (cudaStream_t)stream, "wm_jpeg8_fwd");
}

extern "C" int wm_jpeg8_quantised(const float* x, int64_t x_sb, int64_t x_sc, int64_t x_sh, float* coef,
                                  int B, int H, int W, const wm_jpeg8_params* p, void* stream) {
    if (B == 0) return WM_OK;      // empty work: nothing to validate or launch
    J8Args a{};
    if (int rc = j8_fill(a, x, x_sb, x_sc, x_sh, B, H, W, p, "wm_jpeg8_quantised")) return rc;
    WM_REQUIRE(coef != nullptr && aligned(coef, 32), WM_E_ALIGN, "wm_jpeg8_quantised: coef must be 32-byte aligned");
    a.coef = coef;
    const bool vec = j8_vec_ok(x, x_sb, x_sc, x_sh, W);
    return j8_fwd_any(a, p->variant, p->subsample, vec, true, (cudaStream_t)stream, "wm_jpeg8_quantised");
}

extern "C" int wm_jpeg8_bwd(const float* x, int64_t x_sb, int64_t x_sc, int64_t x_sh,
                            const float* gy, int64_t g_sb, int64_t g_sc, int64_t g_sh, void* gx, int gx_dtype,
                            int B, int H, int W, const wm_jpeg8_params* p, void* stream) {
    if (B == 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(p != nullptr && gx != nullptr && gy != nullptr, WM_E_NULL, "wm_jpeg8_bwd: null pointer");
    WM_REQUIRE(dtype_ok(gx_dtype), WM_E_ARG, "wm_jpeg8_bwd: unknown gx element type %d", gx_dtype);
    cudaStream_t st = (cudaStream_t)stream;
    if (p->variant == WM_JPEG8_HARD) {
        // torch.round has zero gradient everywhere (noise_layers/jpeg.py:233 with round_func=torch.round)
        cudaError_t e = cudaMemsetAsync(gx, 0, dtype_size(gx_dtype) * 3 * size_t(B) * H * W, st);
        return e == cudaSuccess ? WM_OK : cuda_fail(e, "wm_jpeg8_bwd(memset)");
    }
    if (p->variant == WM_JPEG8_MASK) {
        // linear layer: the adjoint is the same pipeline with transposed, swapped colour
        // matrices (the DCT is orthonormal, the mask diagonal) and the subsample adjoint
        wm_jpeg8_params q = *p;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                q.fwd_color[3 * i + j] = p->inv_color[3 * j + i];
                q.inv_color[3 * i + j] = p->fwd_color[3 * j + i];
            }
        q.subsample = 0;
        J8Args a{};
        if (int rc = j8_fill(a, gy, g_sb, g_sc, g_sh, B, H, W, &q, "wm_jpeg8_bwd")) return rc;
        a.out = reinterpret_cast<float*>(gx); a.out_dt = gx_dtype;
        const bool vec = j8_vec_ok(gy, g_sb, g_sc, g_sh, W) && aligned(gx, 8 * dtype_size(gx_dtype));
        J8_TYPED_CHECK(gx_dtype, vec && p->subsample == 0, "wm_jpeg8_bwd");
        return j8_fwd_any(a, WM_JPEG8_MASK, p->subsample == 2 ? 3 : 0, vec, false, st, "wm_jpeg8_bwd");
    }
    J8_TYPED_CHECK(gx_dtype, false, "wm_jpeg8_bwd (JpegSS recompute path: use the saved-state pair)");
    J8Args a{};
    if (int rc = j8_fill(a, x, x_sb, x_sc, x_sh, B, H, W, p, "wm_jpeg8_bwd")) return rc;
    a.gy = gy; a.g_sb = g_sb; a.g_sc = g_sc; a.g_sh = g_sh; a.out = reinterpret_cast<float*>(gx);
    const bool vec = j8_vec_ok(x, x_sb, x_sc, x_sh, W) && j8_vec_ok(gy, g_sb, g_sc, g_sh, W) && aligned(gx, 32);
    const size_t smem = 80 * J8B_THREADS * sizeof(float4);
    if (p->subsample == 2)
        return vec ? j8_launch(jpeg8_ss_bwd_kernel<2, true>, a, J8B_THREADS, smem, st, "wm_jpeg8_bwd")
                   : j8_launch(jpeg8_ss_bwd_kernel<2, false>, a, J8B_THREADS, smem, st, "wm_jpeg8_bwd");
    return vec ? j8_launch(jpeg8_ss_bwd_kernel<0, true>, a, J8B_THREADS, smem, st, "wm_jpeg8_bwd")
               : j8_launch(jpeg8_ss_bwd_kernel<0, false>, a, J8B_THREADS, smem, st, "wm_jpeg8_bwd");
}

// JpegSS training pair (see jpeg8_pair_kernel, DMODE 1 / 2).  Needs the fast-path geometry:
// W % 8 == 0, 32-byte aligned base pointers, strides multiples of 8, subsample == 0.
// d: [B, 3, ceil8(H), W] floats.
extern "C" int wm_jpeg8_fwd_save(const void* x, int x_dtype, int64_t x_sb, int64_t x_sc, int64_t x_sh, float* y, float* d,
                                 int B, int H, int W, const wm_jpeg8_params* p, void* stream) {
    if (B == 0) return WM_OK;      // empty work: nothing to validate or launch
    J8Args a{};
    if (int rc = j8_fill(a, reinterpret_cast<const float*>(x), x_sb, x_sc, x_sh, B, H, W, p, "wm_jpeg8_fwd_save")) return rc;
    WM_REQUIRE(y && d, WM_E_NULL, "wm_jpeg8_fwd_save: null output");
    WM_REQUIRE(dtype_ok(x_dtype), WM_E_ARG, "wm_jpeg8_fwd_save: unknown element type %d", x_dtype);
    WM_REQUIRE(p->variant == WM_JPEG8_SS && p->subsample == 0, WM_E_ARG, "wm_jpeg8_fwd_save: JpegSS without subsampling only");
    WM_REQUIRE(j8_vec_ok(x, x_sb, x_sc, x_sh, W, x_dtype) && aligned(y, 32) && aligned(d, 16), WM_E_ALIGN,
               "wm_jpeg8_fwd_save: needs W %% 8 == 0, x aligned to 8 elements, 32-byte aligned y and strides multiples of 8");
    a.out = y; a.coef = d; a.x_dt = x_dtype;
    return j8_pair_launch<WM_JPEG8_SS, 1>(a, (cudaStream_t)stream, "wm_jpeg8_fwd_save");
}
extern "C" int wm_jpeg8_bwd_saved(const float* gy, int64_t g_sb, int64_t g_sc, int64_t g_sh, const float* d, void* gx, int gx_dtype,
                                  int B, int H, int W, const wm_jpeg8_params* p, void* stream) {
    if (B == 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(p && gy && d && gx, WM_E_NULL, "wm_jpeg8_bwd_saved: null pointer");
    WM_REQUIRE(dtype_ok(gx_dtype), WM_E_ARG, "wm_jpeg8_bwd_saved: unknown gx element type %d", gx_dtype);
    wm_jpeg8_params q = *p;              // adjoint colour matrices, as for the linear variants
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            q.fwd_color[3 * i + j] = p->inv_color[3 * j + i];
            q.inv_color[3 * i + j] = p->fwd_color[3 * j + i];
        }
    q.variant = WM_JPEG8_MASK; q.subsample = 0;
    J8Args a{};
    if (int rc = j8_fill(a, gy, g_sb, g_sc, g_sh, B, H, W, &q, "wm_jpeg8_bwd_saved")) return rc;
    WM_REQUIRE(j8_vec_ok(gy, g_sb, g_sc, g_sh, W) && aligned(gx, 8 * dtype_size(gx_dtype)) && aligned(d, 16), WM_E_ALIGN,
               "wm_jpeg8_bwd_saved: needs W %% 8 == 0, 32-byte aligned gy, gx aligned to 8 elements and strides multiples of 8");
    a.out = reinterpret_cast<float*>(gx); a.out_dt = gx_dtype; a.coef = const_cast<float*>(d);
    return j8_pair_launch<WM_JPEG8_MASK, 2>(a, (cudaStream_t)stream, "wm_jpeg8_bwd_saved");
}
