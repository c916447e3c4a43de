// DiffJPEG forward kernel + C entry (see diffjpeg_core.cuh for the design notes).
#include "diffjpeg_core.cuh"

namespace wm {

// =============================================================================================
// forward
// =============================================================================================
template <int ROUND>
__global__ void __launch_bounds__(DJ_THREADS, 4) diffjpeg_fwd_kernel(const DJArgs a) {
    extern __shared__ float4 smem[];
    float4* scr = smem + threadIdx.x;
    const DJThread t = dj_locate(a);
    const float f = a.factor_ps ? __ldg(a.factor_ps + t.b) : a.factor;

    float cb[4][4], cr[4][4], dummy[4][4];
    dj_load_block(a, t, scr, cb, cr);
    dj_luma_columns<ROUND, false, false>(scr, nullptr, f);
    QuadCoef qx, qy;
    quad_coef_init(qx, t.bx);
    quad_coef_init(qy, t.by);
    dj_chroma_roundtrip<ROUND, false, false>(cb, dummy, qx, qy, t.bx, t.by, f);
    dj_chroma_roundtrip<ROUND, false, false>(cr, dummy, qx, qy, t.bx, t.by, f);

    float* yo = a.out + (int64_t(t.b) * 3 * a.H + t.row0) * a.W + t.col0;
    const int64_t plane = int64_t(a.H) * a.W;
    float tR[4], tG[4], tB[4];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        float yv[8];
        scr_load_row(scr, r, yv);
        idct8(yv);
        if ((r & 1) == 0) dj_chroma_terms(cb[r >> 1], cr[r >> 1], tR, tG, tB);
        f8 oR, oG, oB;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            // min(255, max(0, v)) / 255  (utils/JPEG.py:467-469) == saturate(v / 255)
            oR.v[c] = __saturatef(fmaf(yv[c], DJ_I255, tR[c >> 1]));
            oG.v[c] = __saturatef(fmaf(yv[c], DJ_I255, tG[c >> 1]));
            oB.v[c] = __saturatef(fmaf(yv[c], DJ_I255, tB[c >> 1]));
        }
        if (t.active) {
            float* p = yo + int64_t(r) * a.W;
            stg256(p, oR);
            stg256(p + plane, oG);
            stg256(p + 2 * plane, oB);
        }
    }
}

}  // namespace wm

using namespace wm;

extern "C" int wm_diffjpeg_fwd(const float* x, int64_t x_sb, int64_t x_sc, int64_t x_sh, float* y,
                               int B, int H, int W, float factor, const float* factor_ps,
                               int rounding, void* stream) {
    if (int rc = dj_check(x, x_sb, x_sc, x_sh, B, H, W, "wm_diffjpeg_fwd")) return rc;
    WM_REQUIRE(y != nullptr && aligned(y, 32), WM_E_ALIGN, "wm_diffjpeg_fwd: y must be non-null, 32-byte aligned");
    DJArgs a = dj_args(B, H, W, factor, factor_ps);
    a.x = x; a.x_sb = x_sb; a.x_sc = x_sc; a.x_sh = x_sh; a.out = y;
    const size_t smem = 16 * DJ_THREADS * sizeof(float4);
    DJ_DISPATCH_ROUND(diffjpeg_fwd_kernel, a, smem, (cudaStream_t)stream, "wm_diffjpeg_fwd")
}

