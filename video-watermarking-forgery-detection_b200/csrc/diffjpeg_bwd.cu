// DiffJPEG analytic backward kernel + C entry (see diffjpeg_core.cuh).
#include "diffjpeg_core.cuh"

namespace wm {

// =============================================================================================
// backward (recompute-from-x)
//   g_v = gy * m / 255,  m = 1 inside (0,255), 1/2 at an exact bound (binary torch.min/max tie
//   rule, utils/JPEG.py:467-468), 0 outside;  Mi^T ; chroma 2x2 sum ; G = D g D^T ;
//   G *= round'(q)  (the table*factor of quantise and dequantise cancel) ; D^T G D ;
//   chroma replicate/4 ; M^T ; x255   — the 1/255 and the 255 cancel.
// =============================================================================================
// g * m(u): m = 1 for 0 < u < 1, 1/2 when u sits exactly on a bound, 0 outside
__device__ __forceinline__ float clamp_tie_mask_mul(float g, float u) {
    const bool open_in = u > 0.f && u < 1.f;
    const bool closed_in = u >= 0.f && u <= 1.f;
    return open_in ? g : (closed_in ? 0.5f * g : 0.f);
}

template <int ROUND>
__global__ void __launch_bounds__(DJ_THREADS, 3) diffjpeg_bwd_kernel(const DJArgs a) {
    extern __shared__ float4 smem[];
    float4* scr = smem + threadIdx.x;                         // luminance working block
    float4* dscr = smem + 16 * DJ_THREADS + threadIdx.x;      // round'(q) of the luminance block
    const DJThread t = dj_locate(a);
    const float f = a.factor_ps ? __ldg(a.factor_ps + t.b) : a.factor;

    float cb[4][4], cr[4][4], dcb[4][4], dcr[4][4];
    dj_load_block(a, t, scr, cb, cr);
    dj_luma_columns<ROUND, false, true>(scr, dscr, f);
    QuadCoef qx, qy;
    quad_coef_init(qx, t.bx);
    quad_coef_init(qy, t.by);
    dj_chroma_roundtrip<ROUND, false, true>(cb, dcb, qx, qy, t.bx, t.by, f);
    dj_chroma_roundtrip<ROUND, false, true>(cr, dcr, qx, qy, t.bx, t.by, f);

    // ---- forward tail + masked cotangent, row by row -----------------------------------------
    const float* gr = a.gy + int64_t(t.b) * a.g_sb + int64_t(t.row0) * a.g_sh + t.col0;
    float gcb[4][4], gcr[4][4];
    float tR[4], tG[4], tB[4];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        float yv[8];
        scr_load_row(scr, r, yv);
        idct8(yv);
        if ((r & 1) == 0) dj_chroma_terms(cb[r >> 1], cr[r >> 1], tR, tG, tB);
        f8 gR, gG, gB;
        if (t.active) {
            const float* p = gr + int64_t(r) * a.g_sh;
            gR = ldg256_stream(p);
            gG = ldg256_stream(p + a.g_sc);
            gB = ldg256_stream(p + 2 * a.g_sc);
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) gR.v[c] = gG.v[c] = gB.v[c] = 0.f;
        }
        float gyv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float mr = clamp_tie_mask_mul(gR.v[c], fmaf(yv[c], DJ_I255, tR[c >> 1]));
            const float mg = clamp_tie_mask_mul(gG.v[c], fmaf(yv[c], DJ_I255, tG[c >> 1]));
            const float mb = clamp_tie_mask_mul(gB.v[c], fmaf(yv[c], DJ_I255, tB[c >> 1]));
            gyv[c] = (mr + mg) + mb;                                   // Mi[:,0] = 1
            const float ccb = fmaf(-0.344136f, mg, 1.772f * mb);       // Mi[:,1]
            const float ccr = fmaf(1.402f, mr, -0.714136f * mg);       // Mi[:,2]
            if ((r & 1) == 0 && (c & 1) == 0) { gcb[r >> 1][c >> 1] = ccb; gcr[r >> 1][c >> 1] = ccr; }
            else { gcb[r >> 1][c >> 1] += ccb; gcr[r >> 1][c >> 1] += ccr; }
        }
        dct8(gyv);
        scr_store_row(scr, r, gyv);
    }

    // ---- luminance: column DCT, times round'(q), column IDCT ----------------------------------
#pragma unroll
    for (int cg = 0; cg < 2; ++cg) {
        float v[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float4 t4 = scr[(2 * r + cg) * DJ_THREADS];
            v[r][0] = t4.x; v[r][1] = t4.y; v[r][2] = t4.z; v[r][3] = t4.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            dct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float4 d4 = dscr[(2 * r + cg) * DJ_THREADS];
            v[r][0] *= d4.x; v[r][1] *= d4.y; v[r][2] *= d4.z; v[r][3] *= d4.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            idct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
#pragma unroll
        for (int r = 0; r < 8; ++r)
            scr[(2 * r + cg) * DJ_THREADS] = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
    }

    // ---- chroma: split DCT, times round'(q), split IDCT ---------------------------------------
    quad_dct_rows(gcb, qx, 1); quad_dct_cols(gcb, qy, 16);
    quad_dct_rows(gcr, qx, 1); quad_dct_cols(gcr, qy, 16);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { gcb[i][j] *= dcb[i][j]; gcr[i][j] *= dcr[i][j]; }
    quad_idct_cols(gcb, qy, 16); quad_idct_rows(gcb, qx, 1);
    quad_idct_cols(gcr, qy, 16); quad_idct_rows(gcr, qx, 1);

    // ---- back through the colour transform ----------------------------------------------------
    float* go = a.out + (int64_t(t.b) * 3 * a.H + t.row0) * a.W + t.col0;
    const int64_t plane = int64_t(a.H) * a.W;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        float gv[8];
        scr_load_row(scr, r, gv);
        idct8(gv);
        if ((r & 1) == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {      // avg-pool adjoint (/4) folded in
                const float b4 = gcb[r >> 1][j] * 0.25f, r4 = gcr[r >> 1][j] * 0.25f;
                tR[j] = fmaf(-0.168736f, b4, 0.5f * r4);
                tG[j] = fmaf(-0.331264f, b4, -0.418688f * r4);
                tB[j] = fmaf(0.5f, b4, -0.081312f * r4);
            }
        }
        f8 oR, oG, oB;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            oR.v[c] = fmaf(0.299f, gv[c], tR[c >> 1]);
            oG.v[c] = fmaf(0.587f, gv[c], tG[c >> 1]);
            oB.v[c] = fmaf(0.114f, gv[c], tB[c >> 1]);
        }
        if (t.active) {
            float* p = go + int64_t(r) * a.W;
            stg256(p, oR);
            stg256(p + plane, oG);
            stg256(p + 2 * plane, oB);
        }
    }
}

}  // namespace wm

using namespace wm;

extern "C" int wm_diffjpeg_bwd(const float* x, int64_t x_sb, int64_t x_sc, int64_t x_sh,
                               const float* gy, int64_t g_sb, int64_t g_sc, int64_t g_sh, float* gx,
                               int B, int H, int W, float factor, const float* factor_ps,
                               int rounding, void* stream) {
    if (int rc = dj_check(x, x_sb, x_sc, x_sh, B, H, W, "wm_diffjpeg_bwd(x)")) return rc;
    if (int rc = dj_check(gy, g_sb, g_sc, g_sh, B, H, W, "wm_diffjpeg_bwd(gy)")) return rc;
    WM_REQUIRE(gx != nullptr && aligned(gx, 32), WM_E_ALIGN, "wm_diffjpeg_bwd: gx must be non-null, 32-byte aligned");
    DJArgs a = dj_args(B, H, W, factor, factor_ps);
    a.x = x; a.x_sb = x_sb; a.x_sc = x_sc; a.x_sh = x_sh;
    a.gy = gy; a.g_sb = g_sb; a.g_sc = g_sc; a.g_sh = g_sh; a.out = gx;
    const size_t smem = 32 * DJ_THREADS * sizeof(float4);
    DJ_DISPATCH_ROUND(diffjpeg_bwd_kernel, a, smem, (cudaStream_t)stream, "wm_diffjpeg_bwd")
}

