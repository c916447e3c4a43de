// TMA (cp.async.bulk.tensor) + mbarrier plumbing shared by the plane-stencil kernels
// (blur.cu, median.cu, resize.cu).
//
// The filters of the attack layer use ZERO padding (GaussianBlur: noise_layers/gaussian_blur.py:47,
// MiddleBlur -> kornia MedianBlur).  A TMA tiled load fills out-of-bounds box elements with
// zeros, also for negative start coordinates, so the halo'd tile of a plane arrives in shared
// memory already padded: no per-element bounds test, no address arithmetic in the SM, and the
// whole tile is one outstanding transaction per stage of the ring.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "wm_common.cuh"

namespace wm {

// ---- host: tensor-map encoding through the driver entry point (no -lcuda link dependency) ----
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled tmap_encoder() {
    static PFN_tmapEncodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
    }
    return fn;
}

// Can [planes, H, W] with element strides (sp, sh, 1) be described by a tiled tensor map?
inline bool tmap_ok(const void* base, int64_t sp, int64_t sh, size_t elem) {
    return aligned(base, 16) && (sh * (int64_t)elem) % 16 == 0 && (sp * (int64_t)elem) % 16 == 0 && tmap_encoder() != nullptr;
}

// 3-D map over planes: dim0 = W (innermost), dim1 = H, dim2 = N; box = (box_w, box_h, box_n).
// Returns 0 on success (CUresult otherwise).
inline int tmap_planes(CUtensorMap* m, CUtensorMapDataType dt, size_t elem, const void* base, int N, int H, int W,
                       int64_t sp, int64_t sh, int box_w, int box_h, int box_n = 1) {
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[2] = {(cuuint64_t)(sh * (int64_t)elem), (cuuint64_t)(sp * (int64_t)elem)};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    auto enc = [&]() {
        return (int)tmap_encoder()(m, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    int rc = enc();
    if (rc == (int)CUDA_ERROR_INVALID_CONTEXT) {
        // a thread that has made no runtime call yet (e.g. an autograd worker whose allocations were
        // served from the caching allocator) has no current driver context: bind the primary one
        cudaFree(nullptr);
        rc = enc();
    }
    return rc;
}

// ---- device ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// box (c0.., c1.., c2) of a 3-D tiled map -> shared memory, completion on `bar`
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// ---- planes whose rows are NOT 16-byte aligned (dense W % 4 != 0): no tensor map exists for them, so the same ring
// of halo tiles is filled by every thread of the CTA with 4-byte cp.async (LDGSTS; source size 0 = zero fill outside
// the plane, i.e. the zero padding), one commit group per tile.  The stencil code behind the ring is unchanged; only
// its 128-bit global stores become scalar (the output rows are not 16-byte aligned either).
struct RaggedSrc { const float* x; int64_t sp, sh; };      // plane n at x + n * sp, rows sh apart
template <int NT>
__device__ __forceinline__ void stage_box_cpasync(float* dst, const RaggedSrc& g, int n, int H, int W,
                                                  int x0, int y0, int bw, int bh) {
    const float* plane = g.x + int64_t(n) * g.sp;
    for (int i = threadIdx.x; i < bw * bh; i += NT) {
        const int ly = i / bw, lx = i - ly * bw, gy = y0 + ly, gx = x0 + lx;
        const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
        const float* src = ok ? plane + int64_t(gy) * g.sh + gx : plane;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst + i)), "l"(src), "r"(ok ? 4 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int PENDING>
__device__ __forceinline__ void cpasync_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }
// 4 values at p .. p+3 of an output row that may end inside the group (columns gx .. gx+3 of a W-wide row)
__device__ __forceinline__ void st4_ragged(float* p, const float4 v, int gx, int W) {
    if (gx < W) p[0] = v.x;
    if (gx + 1 < W) p[1] = v.y;
    if (gx + 2 < W) p[2] = v.z;
    if (gx + 3 < W) p[3] = v.w;
}

__device__ __forceinline__ float4 ldg128_nc(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg128(float* p, const float4 v) {
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- shared-memory tiles staged in the caller's element type (DT = WM_DT_*, compile time) ---------------------
// A TMA box of float16 / bfloat16 rows lands as it is; the stencil widens on the way to registers (exact), so an
// autocast trainer's half tensors need no .float() pass.  i = element index in the tile.
template <int DT> __device__ __forceinline__ float tile_widen(uint32_t h16) {
    if (DT == WM_DT_BF16) return __uint_as_float(h16 << 16);
    float f; asm("{\n .reg .b16 t;\n cvt.u16.u32 t, %1;\n cvt.f32.f16 %0, t;\n}" : "=f"(f) : "r"(h16)); return f;
}
template <int DT> __device__ __forceinline__ float2 tile_widen2(uint32_t w) {
    if (DT == WM_DT_BF16) return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
    float lo, hi;
    asm("{\n .reg .b16 l, h;\n mov.b32 {l, h}, %2;\n cvt.f32.f16 %0, l;\n cvt.f32.f16 %1, h;\n}" : "=f"(lo), "=f"(hi) : "r"(w));
    return make_float2(lo, hi);
}
template <int DT> __device__ __forceinline__ float tile_ld1(const void* tile, int i) {
    if (DT == WM_DT_F32) return reinterpret_cast<const float*>(tile)[i];
    return tile_widen<DT>(reinterpret_cast<const uint16_t*>(tile)[i]);
}
template <int DT> __device__ __forceinline__ float2 tile_ld2(const void* tile, int i) {       // i % 2 == 0
    if (DT == WM_DT_F32) return *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(tile) + i);
    return tile_widen2<DT>(*reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(tile) + i));
}
template <int DT> __device__ __forceinline__ float4 tile_ld4(const void* tile, int i) {       // i % 4 == 0
    if (DT == WM_DT_F32) return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(tile) + i);
    const uint2 w = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(tile) + i);
    const float2 a = tile_widen2<DT>(w.x), b = tile_widen2<DT>(w.y);
    return make_float4(a.x, a.y, b.x, b.y);
}
template <int DT> constexpr int tile_elem_size() { return DT == WM_DT_F32 ? 4 : 2; }
template <int DT> constexpr CUtensorMapDataType tile_tmap_type() {
    return DT == WM_DT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : DT == WM_DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
}
// 4 values to element offset `off` (off % 4 == 0) of a float32 / float16 / bfloat16 array, round-to-nearest-even
template <int DT> __device__ __forceinline__ void stg4_typed(void* base, int64_t off, const float4 v) {
    if (DT == WM_DT_F32) { stg128(reinterpret_cast<float*>(base) + off, v); return; }
    uint32_t lo, hi;
    if (DT == WM_DT_BF16) {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(v.y), "f"(v.x));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v.w), "f"(v.z));
    } else {
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(v.y), "f"(v.x));
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v.w), "f"(v.z));
    }
    asm volatile("st.global.v2.u32 [%0], {%1,%2};" ::"l"(reinterpret_cast<uint16_t*>(base) + off), "r"(lo), "r"(hi) : "memory");
}

}  // namespace wm
