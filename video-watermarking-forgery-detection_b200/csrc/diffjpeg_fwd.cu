// DiffJPEG forward kernel + C entry (see diffjpeg_core.cuh for the design notes).
#include "diffjpeg_core.cuh"

namespace wm {

template <int ROUND, bool EP = false, bool TYPED = false>
__global__ void __launch_bounds__(DJ_THREADS, 4) diffjpeg_fwd_kernel(const DJArgs a) {
    extern __shared__ float4 smem[];
    float4* scr = smem + threadIdx.x;
    const DJThread t = dj_locate(a);
    const float f = a.factor_ps ? __ldg(a.factor_ps + t.b) : a.factor;
    dj_load_block<DJ_THREADS, TYPED>(a, t, scr);
    dj_luma_columns<ROUND, false, false, DJ_THREADS>(scr, f);
    QuadCoef qx, qy;
    quad_coef_init(qx, t.bx);
    quad_coef_init(qy, t.by);
    dj_chroma_planes<ROUND, false, false, DJ_THREADS>(scr, qx, qy, t.bx, t.by, f);
    dj_emit_rgb<DJ_THREADS, EP>(a, t, scr);
}

// Forward that also saves what the backward needs (7 B/px: round'(q) of every coefficient and the
// clamp codes), so that wm_diffjpeg_bwd_saved can skip the forward recomputation.
template <int ROUND, bool TYPED = false>
__global__ void __launch_bounds__(DJ_THREADS, 4) diffjpeg_fwd_save_kernel(const DJArgs a) {
    extern __shared__ float4 smem[];
    float4* scr = smem + threadIdx.x;
    const DJThread t = dj_locate(a);
    const float f = a.factor_ps ? __ldg(a.factor_ps + t.b) : a.factor;
    dj_load_block<DJ_THREADS, TYPED>(a, t, scr);
    dj_luma_columns_save<ROUND, DJ_THREADS>(scr, f, a.dY + (int64_t(t.b) * a.H + t.row0) * a.W + t.col0, a.W, t.active);
    QuadCoef qx, qy;
    quad_coef_init(qx, t.bx);
    quad_coef_init(qy, t.by);
    const int64_t Wc = a.W >> 1, plane_c = int64_t(a.H >> 1) * Wc;
    dj_chroma_planes_save<ROUND, DJ_THREADS>(scr, qx, qy, t.bx, t.by, f,
        a.dC + int64_t(t.b) * 2 * plane_c + int64_t(t.mcu_y * 8 + t.by * 4) * Wc + t.mcu_x * 8 + t.bx * 4, plane_c, Wc, t.active);
    dj_emit_rgb_save<DJ_THREADS>(a, t, scr);
}

}  // namespace wm

using namespace wm;

extern "C" int wm_diffjpeg_fwd_save(const void* x, int x_dtype, int64_t x_sb, int64_t x_sc, int64_t x_sh, float* y,
                                    float* dY, float* dC, uint64_t* clamp_codes, int B, int H, int W,
                                    float factor, const float* factor_ps, int rounding, void* stream) {
    if (B == 0) return WM_OK;      // empty work: nothing to validate or launch
    if (int rc = dj_check(x, x_sb, x_sc, x_sh, B, H, W, "wm_diffjpeg_fwd_save", x_dtype)) return rc;
    WM_REQUIRE(y && dY && dC && clamp_codes, WM_E_NULL, "wm_diffjpeg_fwd_save: null output pointer");
    WM_REQUIRE(aligned(y, 32) && aligned(dY, 16) && aligned(dC, 16) && aligned(clamp_codes, 8), WM_E_ALIGN,
               "wm_diffjpeg_fwd_save: y must be 32-byte, dY/dC 16-byte, clamp_codes 8-byte aligned");
    DJArgs a = dj_args(B, H, W, factor, factor_ps);
    a.x = x; a.x_dt = x_dtype; a.x_sb = x_sb; a.x_sc = x_sc; a.x_sh = x_sh; a.out = y;
    a.dY = dY; a.dC = dC; a.cm = reinterpret_cast<unsigned long long*>(clamp_codes);
    const size_t smem = SC_FWD_CHUNKS * DJ_THREADS * sizeof(float4);
    if (x_dtype != WM_DT_F32) { DJ_DISPATCH_ROUND_T(diffjpeg_fwd_save_kernel, true, a, DJ_THREADS, smem, (cudaStream_t)stream, "wm_diffjpeg_fwd_save") }
    DJ_DISPATCH_ROUND_T(diffjpeg_fwd_save_kernel, false, a, DJ_THREADS, smem, (cudaStream_t)stream, "wm_diffjpeg_fwd_save")
}


extern "C" int wm_diffjpeg_fwd(const void* x, int x_dtype, int64_t x_sb, int64_t x_sc, int64_t x_sh, float* y,
                               int B, int H, int W, float factor, const float* factor_ps,
                               int rounding, const wm_store_epilogue* ep, void* stream) {
    if (B == 0) return WM_OK;      // empty work: nothing to validate or launch
    if (int rc = dj_check(x, x_sb, x_sc, x_sh, B, H, W, "wm_diffjpeg_fwd", x_dtype)) return rc;
    WM_REQUIRE(y != nullptr && aligned(y, 32), WM_E_ALIGN, "wm_diffjpeg_fwd: y must be non-null, 32-byte aligned");
    DJArgs a = dj_args(B, H, W, factor, factor_ps);
    a.x = x; a.x_dt = x_dtype; a.x_sb = x_sb; a.x_sc = x_sc; a.x_sh = x_sh; a.out = y;
    const size_t smem = SC_FWD_CHUNKS * DJ_THREADS * sizeof(float4);
    WM_EP_CHECK(ep, "wm_diffjpeg_fwd");
    a.ep = make_store_ep(ep);
    WM_REQUIRE(!a.ep.x || x_dtype == WM_DT_F32, WM_E_ARG, "wm_diffjpeg_fwd: the store epilogue needs a float32 image");
    if (x_dtype != WM_DT_F32) {
        switch (rounding) {
            case WM_ROUND_ONLY_AT_0: return dj_launch(diffjpeg_fwd_kernel<WM_ROUND_ONLY_AT_0, false, true>, a, DJ_THREADS, smem, (cudaStream_t)stream, "wm_diffjpeg_fwd");
            case WM_ROUND_CUBIC:     return dj_launch(diffjpeg_fwd_kernel<WM_ROUND_CUBIC, false, true>, a, DJ_THREADS, smem, (cudaStream_t)stream, "wm_diffjpeg_fwd");
            case WM_ROUND_HARD:      return dj_launch(diffjpeg_fwd_kernel<WM_ROUND_HARD, false, true>, a, DJ_THREADS, smem, (cudaStream_t)stream, "wm_diffjpeg_fwd");
            case WM_ROUND_FOURIER:   return dj_launch(diffjpeg_fwd_kernel<WM_ROUND_FOURIER, false, true>, a, DJ_THREADS, smem, (cudaStream_t)stream, "wm_diffjpeg_fwd");
            default: set_error("unknown rounding mode %d", rounding); return WM_E_ARG;
        }
    }
    if (a.ep.x) {
        switch (rounding) {
            case WM_ROUND_ONLY_AT_0: return dj_launch(diffjpeg_fwd_kernel<WM_ROUND_ONLY_AT_0, true>, a, DJ_THREADS, smem, (cudaStream_t)stream, "wm_diffjpeg_fwd");
            case WM_ROUND_CUBIC:     return dj_launch(diffjpeg_fwd_kernel<WM_ROUND_CUBIC, true>, a, DJ_THREADS, smem, (cudaStream_t)stream, "wm_diffjpeg_fwd");
            case WM_ROUND_HARD:      return dj_launch(diffjpeg_fwd_kernel<WM_ROUND_HARD, true>, a, DJ_THREADS, smem, (cudaStream_t)stream, "wm_diffjpeg_fwd");
            case WM_ROUND_FOURIER:   return dj_launch(diffjpeg_fwd_kernel<WM_ROUND_FOURIER, true>, a, DJ_THREADS, smem, (cudaStream_t)stream, "wm_diffjpeg_fwd");
            default: set_error("unknown rounding mode %d", rounding); return WM_E_ARG;
        }
    }
    DJ_DISPATCH_ROUND(diffjpeg_fwd_kernel, a, DJ_THREADS, smem, (cudaStream_t)stream, "wm_diffjpeg_fwd")
}
