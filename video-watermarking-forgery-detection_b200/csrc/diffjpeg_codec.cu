// DiffJPEG compress / decompress kernels + C entries (see diffjpeg_core.cuh).
#include "diffjpeg_core.cuh"

namespace wm {

// =============================================================================================
// compress: rounded quantised coefficients in the reference's [B, nblk, 8, 8] layout
// =============================================================================================
template <int ROUND>
__global__ void __launch_bounds__(DJ_THREADS, 4) diffjpeg_compress_kernel(const DJArgs a) {
    extern __shared__ float4 smem[];
    float4* scr = smem + threadIdx.x;
    const DJThread t = dj_locate(a);
    const float f = a.factor_ps ? __ldg(a.factor_ps + t.b) : a.factor;
    dj_load_block<DJ_THREADS>(a, t, scr);
    dj_luma_columns<ROUND, true, false, DJ_THREADS>(scr, f);
    QuadCoef qx, qy;
    quad_coef_init(qx, t.bx);
    quad_coef_init(qy, t.by);
    dj_chroma_planes<ROUND, true, false, DJ_THREADS>(scr, qx, qy, t.bx, t.by, f);
    if (!t.active) return;
    // luminance block index (utils/JPEG.py:176-181): raster over (H/8, W/8)
    const int64_t yblk = (int64_t(t.b) * (a.H / 8) + t.row0 / 8) * (a.W / 8) + t.col0 / 8;
    float* py = a.coef_y + yblk * 64;
#pragma unroll 1
    for (int u = 0; u < 8; ++u) {
        float v[8];
        scr_load_row<DJ_THREADS>(scr, u, v);
        f8 o;
#pragma unroll
        for (int c = 0; c < 8; ++c) o.v[c] = v[c];
        stg256(py + u * 8, o);
    }
    const int64_t cblk = (int64_t(t.b) * (a.H / 16) + t.mcu_y) * (a.W / 16) + t.mcu_x;
#pragma unroll 1
    for (int pl = 0; pl < 2; ++pl) {
        float* pc = (pl ? a.coef_cr : a.coef_cb) + cblk * 64;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v[4];
            f4_to(v, scr[(SC_CB + 4 * pl + i) * DJ_THREADS]);
#pragma unroll
            for (int j = 0; j < 4; ++j) pc[(2 * i + t.by) * 8 + 2 * j + t.bx] = v[j];
        }
    }
}

// =============================================================================================
// decompress: coefficients -> image (utils/JPEG.py:452-469)
// =============================================================================================
__global__ void __launch_bounds__(DJ_THREADS, 4) diffjpeg_decompress_kernel(const DJArgs a) {
    constexpr int NT = DJ_THREADS;
    extern __shared__ float4 smem[];
    float4* scr = smem + threadIdx.x;
    const DJThread t = dj_locate(a);
    const float f = a.factor_ps ? __ldg(a.factor_ps + t.b) : a.factor;
    const int64_t yblk = (int64_t(t.b) * (a.H / 8) + t.row0 / 8) * (a.W / 8) + t.col0 / 8;
    const int64_t cblk = (int64_t(t.b) * (a.H / 16) + t.mcu_y) * (a.W / 16) + t.mcu_x;
    // luminance: dequantise, column IDCT per 4-column group via scratch, then row IDCT
#pragma unroll 1
    for (int u = 0; u < 8; ++u) {
        float v[8];
        if (t.active) {
            f8 q = ldg256_stream(a.coef_y + yblk * 64 + u * 8);
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = q.v[c] * (cTY[u * 8 + c] * f);
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = 0.f;
        }
        scr_store_row<NT>(scr, u, v);
    }
#pragma unroll 1
    for (int cg = 0; cg < 2; ++cg) {
        float v[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r) f4_to(v[r], scr[(SC_Y + 2 * r + cg) * NT]);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            idct8(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
#pragma unroll
        for (int r = 0; r < 8; ++r) scr[(SC_Y + 2 * r + cg) * NT] = to_f4(v[r]);
    }
    QuadCoef qx, qy;
    quad_coef_init(qx, t.bx);
    quad_coef_init(qy, t.by);
#pragma unroll 1
    for (int pl = 0; pl < 2; ++pl) {
        const float* pc = (pl ? a.coef_cr : a.coef_cb) + cblk * 64;
        float p[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int o = (2 * i + t.by) * 8 + 2 * j + t.bx;
                p[i][j] = t.active ? __ldg(pc + o) * (cTC[o] * f) : 0.f;
            }
        quad_idct_cols(p, qy, 16);
        quad_idct_rows(p, qx, 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) scr[(SC_CB + 4 * pl + i) * NT] = to_f4(p[i]);
    }
    dj_emit_rgb<NT>(a, t, scr);
}

}  // namespace wm

using namespace wm;

extern "C" int wm_diffjpeg_compress(const float* x, int64_t x_sb, int64_t x_sc, int64_t x_sh,
                                    float* coef_y, float* coef_cb, float* coef_cr, int B, int H, int W,
                                    float factor, const float* factor_ps, int rounding, void* stream) {
    if (B == 0) return WM_OK;      // empty work: nothing to validate or launch
    if (int rc = dj_check(x, x_sb, x_sc, x_sh, B, H, W, "wm_diffjpeg_compress")) return rc;
    WM_REQUIRE(coef_y && coef_cb && coef_cr && aligned(coef_y, 32), WM_E_NULL,
               "wm_diffjpeg_compress: coefficient outputs must be non-null (coef_y 32-byte aligned)");
    DJArgs a = dj_args(B, H, W, factor, factor_ps);
    a.x = x; a.x_dt = WM_DT_F32; a.x_sb = x_sb; a.x_sc = x_sc; a.x_sh = x_sh;
    a.coef_y = coef_y; a.coef_cb = coef_cb; a.coef_cr = coef_cr;
    const size_t smem = SC_FWD_CHUNKS * DJ_THREADS * sizeof(float4);
    DJ_DISPATCH_ROUND(diffjpeg_compress_kernel, a, DJ_THREADS, smem, (cudaStream_t)stream, "wm_diffjpeg_compress")
}

extern "C" int wm_diffjpeg_decompress(const float* coef_y, const float* coef_cb, const float* coef_cr,
                                      float* y, int B, int H, int W, float factor,
                                      const float* factor_ps, void* stream) {
    if (B == 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(coef_y && coef_cb && coef_cr && y, WM_E_NULL, "wm_diffjpeg_decompress: null pointer");
    WM_REQUIRE(B >= 0 && H > 0 && W > 0 && H % 16 == 0 && W % 16 == 0, WM_E_SHAPE,
               "wm_diffjpeg_decompress: H and W must be positive multiples of 16 (got %d x %d)", H, W);
    WM_REQUIRE(aligned(coef_y, 32) && aligned(y, 32), WM_E_ALIGN, "wm_diffjpeg_decompress: 32-byte alignment required");
    DJArgs a = dj_args(B, H, W, factor, factor_ps);
    a.coef_y = const_cast<float*>(coef_y); a.coef_cb = const_cast<float*>(coef_cb);
    a.coef_cr = const_cast<float*>(coef_cr); a.out = y;
    const size_t smem = SC_FWD_CHUNKS * DJ_THREADS * sizeof(float4);
    return dj_launch(diffjpeg_decompress_kernel, a, DJ_THREADS, smem, (cudaStream_t)stream, "wm_diffjpeg_decompress");
}
