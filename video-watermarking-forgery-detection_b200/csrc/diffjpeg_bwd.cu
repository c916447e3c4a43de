// DiffJPEG analytic backward kernel + C entry (see diffjpeg_core.cuh).
//
//   g_v = gy * m / 255,  m = 1 inside (0,255), 1/2 at an exact bound (binary torch.min/max tie
//   rule, utils/JPEG.py:467-468), 0 outside;  Mi^T ; chroma 2x2 sum ; G = D g D^T ;
//   G *= round'(q)  (the table*factor of quantise and dequantise cancel) ; D^T G D ;
//   chroma replicate/4 ; M^T ; x255   — the 1/255 and the 255 cancel.
// Everything the chain needs (q for round', the pre-clamp RGB for m) is RECOMPUTED from x in the
// same kernel: HBM traffic is read x + read gy + write gx = 36 B/px and the forward saves nothing.
#include "diffjpeg_core.cuh"

namespace wm {

// g * m(u): m = 1 for 0 < u < 1, 1/2 when u sits exactly on a bound, 0 outside
__device__ __forceinline__ float clamp_tie_mask_mul(float g, float u) {
    const bool open_in = u > 0.f && u < 1.f;
    const bool closed_in = u >= 0.f && u <= 1.f;
    return open_in ? g : (closed_in ? 0.5f * g : 0.f);
}

template <int ROUND, bool TYPED = false>
__global__ void __launch_bounds__(DJB_THREADS, 3) diffjpeg_bwd_kernel(const DJArgs a) {
    constexpr int NT = DJB_THREADS;
    const int odt = TYPED ? a.out_dt : WM_DT_F32;
    extern __shared__ float4 smem[];
    float4* scr = smem + threadIdx.x;
    const DJThread t = dj_locate(a);
    const float f = a.factor_ps ? __ldg(a.factor_ps + t.b) : a.factor;
    const float* gr = a.gy + int64_t(t.b) * a.g_sb + int64_t(t.row0) * a.g_sh + t.col0;
    if (t.active) {           // the cotangent is needed ~2/3 into the kernel: pull it into L2 now
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            prefetch_l2(gr + int64_t(r) * a.g_sh);
            prefetch_l2(gr + int64_t(r) * a.g_sh + a.g_sc);
            prefetch_l2(gr + int64_t(r) * a.g_sh + 2 * a.g_sc);
        }
    }

    // ---- forward recompute: Y -> SC_Y (column-inverse-transformed), chroma -> SC_CB/SC_CR,
    //      round'(q) -> SC_DY / SC_DC
    dj_load_block<NT, TYPED>(a, t, scr);
    dj_luma_columns<ROUND, false, true, NT>(scr, f);
    QuadCoef qx, qy;
    quad_coef_init(qx, t.bx);
    quad_coef_init(qy, t.by);
    dj_chroma_planes<ROUND, false, true, NT>(scr, qx, qy, t.bx, t.by, f);

    // ---- forward tail + masked cotangent, one row pair per iteration ---------------------------
#pragma unroll 1
    for (int rp = 0; rp < 4; ++rp) {
        RowPair g;
        dj_load_pair(g, gr, int64_t(2 * rp) * a.g_sh, a.g_sh, a.g_sc, t.active);
        float cb[4], cr[4], tR[4], tG[4], tB[4], gcb[4], gcr[4];
        f4_to(cb, scr[(SC_CB + rp) * NT]);
        f4_to(cr, scr[(SC_CR + rp) * NT]);
        dj_chroma_terms(cb, cr, tR, tG, tB);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int r = 2 * rp + rr;
            float yv[8], gyv[8];
            scr_load_row<NT>(scr, r, yv);
            idct8(yv);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float mr = clamp_tie_mask_mul(g.R[rr].v[c], fmaf(yv[c], DJ_I255, tR[c >> 1]));
                const float mg = clamp_tie_mask_mul(g.G[rr].v[c], fmaf(yv[c], DJ_I255, tG[c >> 1]));
                const float mb = clamp_tie_mask_mul(g.B[rr].v[c], fmaf(yv[c], DJ_I255, tB[c >> 1]));
                gyv[c] = (mr + mg) + mb;                                   // Mi[:,0] = 1
                const float ccb = fmaf(-0.344136f, mg, 1.772f * mb);       // Mi[:,1]
                const float ccr = fmaf(1.402f, mr, -0.714136f * mg);       // Mi[:,2]
                if (rr == 0 && (c & 1) == 0) { gcb[c >> 1] = ccb; gcr[c >> 1] = ccr; }
                else { gcb[c >> 1] += ccb; gcr[c >> 1] += ccr; }
            }
            dct8(gyv);
            scr_store_row<NT>(scr, r, gyv);
        }
        scr[(SC_CB + rp) * NT] = to_f4(gcb);       // the forward chroma of this pair is consumed
        scr[(SC_CR + rp) * NT] = to_f4(gcr);
    }

    // ---- luminance: column DCT, times round'(q), column IDCT ----------------------------------
#pragma unroll 1
    for (int cg = 0; cg < 2; ++cg) {
        float v[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r) f4_to(v[r], scr[(SC_Y + 2 * r + cg) * NT]);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            dct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float4 d4 = scr[(SC_DY + 2 * r + cg) * NT];
            v[r][0] *= d4.x; v[r][1] *= d4.y; v[r][2] *= d4.z; v[r][3] *= d4.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            idct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
#pragma unroll
        for (int r = 0; r < 8; ++r) scr[(SC_Y + 2 * r + cg) * NT] = to_f4(v[r]);
    }

    // ---- chroma: split DCT, times round'(q), split IDCT ---------------------------------------
#pragma unroll 1
    for (int pl = 0; pl < 2; ++pl) {
        float p[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) f4_to(p[i], scr[(SC_CB + 4 * pl + i) * NT]);
        quad_dct_rows(p, qx, 1);
        quad_dct_cols(p, qy, 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 d4 = scr[(SC_DC + 4 * pl + i) * NT];
            p[i][0] *= d4.x; p[i][1] *= d4.y; p[i][2] *= d4.z; p[i][3] *= d4.w;
        }
        quad_idct_cols(p, qy, 16);
        quad_idct_rows(p, qx, 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) scr[(SC_CB + 4 * pl + i) * NT] = to_f4(p[i]);
    }

    // ---- back through the colour transform ----------------------------------------------------
    const int64_t go = (int64_t(t.b) * 3 * a.H + t.row0) * a.W + t.col0;       // element offset into gx (a.out_dt elements)
    const int64_t plane = int64_t(a.H) * a.W;
#pragma unroll 1
    for (int rp = 0; rp < 4; ++rp) {
        float gcb[4], gcr[4], tR[4], tG[4], tB[4];
        f4_to(gcb, scr[(SC_CB + rp) * NT]);
        f4_to(gcr, scr[(SC_CR + rp) * NT]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {      // avg-pool adjoint (/4) folded in
            const float b4 = gcb[j] * 0.25f, r4 = gcr[j] * 0.25f;
            tR[j] = fmaf(-0.168736f, b4, 0.5f * r4);
            tG[j] = fmaf(-0.331264f, b4, -0.418688f * r4);
            tB[j] = fmaf(0.5f, b4, -0.081312f * r4);
        }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int r = 2 * rp + rr;
            float gv[8];
            scr_load_row<NT>(scr, r, gv);
            idct8(gv);
            f8 oR, oG, oB;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                oR.v[c] = fmaf(0.299f, gv[c], tR[c >> 1]);
                oG.v[c] = fmaf(0.587f, gv[c], tG[c >> 1]);
                oB.v[c] = fmaf(0.114f, gv[c], tB[c >> 1]);
            }
            if (t.active) {
                const int64_t p = go + int64_t(r) * a.W;
                st8_typed(a.out, p, oR, odt);
                st8_typed(a.out, p + plane, oG, odt);
                st8_typed(a.out, p + 2 * plane, oB, odt);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Backward from the state saved by wm_diffjpeg_fwd_save: gy, round'(q) and the clamp codes in,
// gx out (31 B/px); no forward recomputation, the forward's 24-chunk scratch and occupancy.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float code_mul(float g, unsigned code) {      // 0 -> 0, 1 -> g, 2 -> g/2
    return (code & 1u) ? g : ((code & 2u) ? 0.5f * g : 0.f);
}

template <bool TYPED = false>
__global__ void __launch_bounds__(DJ_THREADS, 4) diffjpeg_bwd_saved_kernel(const DJArgs a) {
    constexpr int NT = DJ_THREADS;
    const int odt = TYPED ? a.out_dt : WM_DT_F32;
    extern __shared__ float4 smem[];
    float4* scr = smem + threadIdx.x;
    const DJThread t = dj_locate(a);
    const float* gr = a.gy + int64_t(t.b) * a.g_sb + int64_t(t.row0) * a.g_sh + t.col0;
    const unsigned long long* cmr = a.cm + (int64_t(t.b) * a.H + t.row0) * (a.W >> 3) + (t.col0 >> 3);
    const float* dYr = a.dY + (int64_t(t.b) * a.H + t.row0) * a.W + t.col0;
    const int64_t Wc = a.W >> 1, plane_c = int64_t(a.H >> 1) * Wc;
    const float* dCr = a.dC + int64_t(t.b) * 2 * plane_c + int64_t(t.mcu_y * 8 + t.by * 4) * Wc + t.mcu_x * 8 + t.bx * 4;

    // ---- masked cotangent through Mi^T, chroma 2x2 sum, row DCT; one row pair per (rolled) iteration ---------------------------------
    auto consume = [&](const RowPair& g, int rp) {
        float gcb[4], gcr[4];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int r = 2 * rp + rr;
            const unsigned long long cw = t.active ? __ldg(cmr + int64_t(r) * (a.W >> 3)) : 0ull;
            const unsigned lo = (unsigned)cw, hi = (unsigned)(cw >> 32);
            float gyv[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const unsigned code = ((c < 4 ? lo : hi) >> (6 * (c & 3))) & 63u;
                const float mr = code_mul(g.R[rr].v[c], code), mg = code_mul(g.G[rr].v[c], code >> 2),
                            mb = code_mul(g.B[rr].v[c], code >> 4);
                gyv[c] = (mr + mg) + mb;                                   // Mi[:,0] = 1
                const float ccb = fmaf(-0.344136f, mg, 1.772f * mb);       // Mi[:,1]
                const float ccr = fmaf(1.402f, mr, -0.714136f * mg);       // Mi[:,2]
                if (rr == 0 && (c & 1) == 0) { gcb[c >> 1] = ccb; gcr[c >> 1] = ccr; }
                else { gcb[c >> 1] += ccb; gcr[c >> 1] += ccr; }
            }
            dct8(gyv);
            scr_store_row<NT>(scr, r, gyv);
        }
        scr[(SC_CB + rp) * NT] = to_f4(gcb);
        scr[(SC_CR + rp) * NT] = to_f4(gcr);
    };
    {
        RowPair A, Bp;          // one pair of loads in flight ahead of the arithmetic
        dj_load_pair(A, gr, 0, a.g_sh, a.g_sc, t.active);
#pragma unroll 1
        for (int rp = 0; rp < 4; ++rp) {
            if (rp < 3) dj_load_pair(Bp, gr, int64_t(2 * rp + 2) * a.g_sh, a.g_sh, a.g_sc, t.active);
            consume(A, rp);
            A = Bp;
        }
    }

    // ---- luminance: column DCT, times round'(q), column IDCT ----------------------------------
#pragma unroll 1
    for (int cg = 0; cg < 2; ++cg) {
        float v[8][4];
        float4 d4[8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
            d4[r] = t.active ? ldg128_stream(dYr + int64_t(r) * a.W + 4 * cg) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 8; ++r) f4_to(v[r], scr[(SC_Y + 2 * r + cg) * NT]);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            dct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
#pragma unroll
        for (int r = 0; r < 8; ++r) { v[r][0] *= d4[r].x; v[r][1] *= d4[r].y; v[r][2] *= d4[r].z; v[r][3] *= d4[r].w; }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            idct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
#pragma unroll
        for (int r = 0; r < 8; ++r) scr[(SC_Y + 2 * r + cg) * NT] = to_f4(v[r]);
    }

    // ---- chroma: split DCT, times round'(q), split IDCT ---------------------------------------
    QuadCoef qx, qy;
    quad_coef_init(qx, t.bx);
    quad_coef_init(qy, t.by);
#pragma unroll 1
    for (int pl = 0; pl < 2; ++pl) {
        float p[4][4];
        float4 d4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            d4[i] = t.active ? ldg128_stream(dCr + pl * plane_c + int64_t(i) * Wc) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) f4_to(p[i], scr[(SC_CB + 4 * pl + i) * NT]);
        quad_dct_rows(p, qx, 1);
        quad_dct_cols(p, qy, 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) { p[i][0] *= d4[i].x; p[i][1] *= d4[i].y; p[i][2] *= d4[i].z; p[i][3] *= d4[i].w; }
        quad_idct_cols(p, qy, 16);
        quad_idct_rows(p, qx, 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) scr[(SC_CB + 4 * pl + i) * NT] = to_f4(p[i]);
    }

    // ---- back through the colour transform ----------------------------------------------------
    const int64_t go = (int64_t(t.b) * 3 * a.H + t.row0) * a.W + t.col0;       // element offset into gx (a.out_dt elements)
    const int64_t plane = int64_t(a.H) * a.W;
#pragma unroll 1
    for (int rp = 0; rp < 4; ++rp) {
        float gcb[4], gcr[4], tR[4], tG[4], tB[4];
        f4_to(gcb, scr[(SC_CB + rp) * NT]);
        f4_to(gcr, scr[(SC_CR + rp) * NT]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {      // avg-pool adjoint (/4) folded in
            const float b4 = gcb[j] * 0.25f, r4 = gcr[j] * 0.25f;
            tR[j] = fmaf(-0.168736f, b4, 0.5f * r4);
            tG[j] = fmaf(-0.331264f, b4, -0.418688f * r4);
            tB[j] = fmaf(0.5f, b4, -0.081312f * r4);
        }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int r = 2 * rp + rr;
            float gv[8];
            scr_load_row<NT>(scr, r, gv);
            idct8(gv);
            f8 oR, oG, oB;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                oR.v[c] = fmaf(0.299f, gv[c], tR[c >> 1]);
                oG.v[c] = fmaf(0.587f, gv[c], tG[c >> 1]);
                oB.v[c] = fmaf(0.114f, gv[c], tB[c >> 1]);
            }
            if (t.active) {
                const int64_t p = go + int64_t(r) * a.W;
                st8_typed(a.out, p, oR, odt);
                st8_typed(a.out, p + plane, oG, odt);
                st8_typed(a.out, p + 2 * plane, oB, odt);
            }
        }
    }
}

}  // namespace wm

using namespace wm;

extern "C" int wm_diffjpeg_bwd_saved(const float* gy, int64_t g_sb, int64_t g_sc, int64_t g_sh,
                                     const float* dY, const float* dC, const uint64_t* clamp_codes, void* gx, int gx_dtype,
                                     int B, int H, int W, void* stream) {
    if (B == 0) return WM_OK;      // empty work: nothing to validate or launch
    if (int rc = dj_check(gy, g_sb, g_sc, g_sh, B, H, W, "wm_diffjpeg_bwd_saved(gy)")) return rc;
    WM_REQUIRE(dY && dC && clamp_codes && gx, WM_E_NULL, "wm_diffjpeg_bwd_saved: null pointer");
    WM_REQUIRE(dtype_ok(gx_dtype), WM_E_ARG, "wm_diffjpeg_bwd_saved: unknown gx element type %d", gx_dtype);
    WM_REQUIRE(aligned(gx, 8 * dtype_size(gx_dtype)) && aligned(dY, 16) && aligned(dC, 16) && aligned(clamp_codes, 8), WM_E_ALIGN,
               "wm_diffjpeg_bwd_saved: gx must be aligned to 8 elements, dY/dC 16-byte, clamp_codes 8-byte aligned");
    DJArgs a = dj_args(B, H, W, 1.f, nullptr);
    a.gy = gy; a.g_sb = g_sb; a.g_sc = g_sc; a.g_sh = g_sh; a.out = reinterpret_cast<float*>(gx); a.out_dt = gx_dtype;
    a.dY = const_cast<float*>(dY); a.dC = const_cast<float*>(dC);
    a.cm = reinterpret_cast<unsigned long long*>(const_cast<uint64_t*>(clamp_codes));
    const size_t smem = SC_FWD_CHUNKS * DJ_THREADS * sizeof(float4);
    if (gx_dtype != WM_DT_F32) return dj_launch(diffjpeg_bwd_saved_kernel<true>, a, DJ_THREADS, smem, (cudaStream_t)stream, "wm_diffjpeg_bwd_saved");
    return dj_launch(diffjpeg_bwd_saved_kernel<false>, a, DJ_THREADS, smem, (cudaStream_t)stream, "wm_diffjpeg_bwd_saved");
}


extern "C" int wm_diffjpeg_bwd(const void* x, int x_dtype, int64_t x_sb, int64_t x_sc, int64_t x_sh,
                               const float* gy, int64_t g_sb, int64_t g_sc, int64_t g_sh, void* gx, int gx_dtype,
                               int B, int H, int W, float factor, const float* factor_ps,
                               int rounding, void* stream) {
    if (B == 0) return WM_OK;      // empty work: nothing to validate or launch
    if (int rc = dj_check(x, x_sb, x_sc, x_sh, B, H, W, "wm_diffjpeg_bwd(x)", x_dtype)) return rc;
    if (int rc = dj_check(gy, g_sb, g_sc, g_sh, B, H, W, "wm_diffjpeg_bwd(gy)")) return rc;
    WM_REQUIRE(dtype_ok(gx_dtype), WM_E_ARG, "wm_diffjpeg_bwd: unknown gx element type %d", gx_dtype);
    WM_REQUIRE(gx != nullptr && aligned(gx, 8 * dtype_size(gx_dtype)), WM_E_ALIGN, "wm_diffjpeg_bwd: gx must be non-null, aligned to 8 elements");
    DJArgs a = dj_args(B, H, W, factor, factor_ps);
    a.x = x; a.x_dt = x_dtype; a.x_sb = x_sb; a.x_sc = x_sc; a.x_sh = x_sh;
    a.gy = gy; a.g_sb = g_sb; a.g_sc = g_sc; a.g_sh = g_sh; a.out = reinterpret_cast<float*>(gx); a.out_dt = gx_dtype;
    const size_t smem = SC_BWD_CHUNKS * DJB_THREADS * sizeof(float4);
    if (x_dtype != WM_DT_F32 || gx_dtype != WM_DT_F32) { DJ_DISPATCH_ROUND_T(diffjpeg_bwd_kernel, true, a, DJB_THREADS, smem, (cudaStream_t)stream, "wm_diffjpeg_bwd") }
    DJ_DISPATCH_ROUND_T(diffjpeg_bwd_kernel, false, a, DJB_THREADS, smem, (cudaStream_t)stream, "wm_diffjpeg_bwd")
}
