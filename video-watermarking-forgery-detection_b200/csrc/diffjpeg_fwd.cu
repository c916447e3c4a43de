// DiffJPEG forward kernel + C entry (see diffjpeg_core.cuh for the design notes).
#include "diffjpeg_core.cuh"

namespace wm {

template <int ROUND>
__global__ void __launch_bounds__(DJ_THREADS, 4) diffjpeg_fwd_kernel(const DJArgs a) {
    extern __shared__ float4 smem[];
    float4* scr = smem + threadIdx.x;
    const DJThread t = dj_locate(a);
    const float f = a.factor_ps ? __ldg(a.factor_ps + t.b) : a.factor;
    dj_load_block<DJ_THREADS>(a, t, scr);
    dj_luma_columns<ROUND, false, false, DJ_THREADS>(scr, f);
    QuadCoef qx, qy;
    quad_coef_init(qx, t.bx);
    quad_coef_init(qy, t.by);
    dj_chroma_planes<ROUND, false, false, DJ_THREADS>(scr, qx, qy, t.bx, t.by, f);
    dj_emit_rgb<DJ_THREADS>(a, t, scr);
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
    const size_t smem = SC_FWD_CHUNKS * DJ_THREADS * sizeof(float4);
    DJ_DISPATCH_ROUND(diffjpeg_fwd_kernel, a, DJ_THREADS, smem, (cudaStream_t)stream, "wm_diffjpeg_fwd")
}
