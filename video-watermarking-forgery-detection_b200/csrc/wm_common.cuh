// Shared device/host helpers for libwmattack (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/wm_attack.h"

namespace wm {

// ---- error plumbing: no C++ exception crosses the C ABI -----------------------------------
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);

#define WM_REQUIRE(cond, code, ...)                         \
    do {                                                    \
        if (!(cond)) { ::wm::set_error(__VA_ARGS__); return (code); } \
    } while (0)

#define WM_LAUNCH_CHECK(what)                               \
    do {                                                    \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return ::wm::cuda_fail(e__, what); \
    } while (0)

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ---- 256-bit global access (sm_100: LDG.E.256 / STG.E.256) --------------------------------
struct __align__(32) f8 { float v[8]; };

__device__ __forceinline__ f8 ldg256_stream(const float* p) {
    f8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]),
                   "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg256(float* p, const f8& r) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p),
                 "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]),
                 "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7])
                 : "memory");
}
__device__ __forceinline__ float4 ldg128_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// ---- typed rows (boundary cast fused into the load / store: SURVEY 8b, autocast at models/IRNcrop_model.py:340) ------
// 8 consecutive elements at element offset `off` of a float32 / float16 / bfloat16 array, as floats (f16 / bf16 -> f32
// is exact), and the reverse with round-to-nearest-even (what torch's .to(dtype) does).  dt is warp-uniform.
__device__ __forceinline__ f8 ld8_typed(const void* base, int64_t off, int dt) {
    if (dt == WM_DT_F32) return ldg256_stream(reinterpret_cast<const float*>(base) + off);
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(reinterpret_cast<const uint16_t*>(base) + off));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    f8 o;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (dt == WM_DT_BF16) {
            o.v[2 * i] = __uint_as_float(w[i] << 16);
            o.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        } else {
            float lo, hi;
            asm("{\n .reg .b16 l, h;\n mov.b32 {l, h}, %2;\n cvt.f32.f16 %0, l;\n cvt.f32.f16 %1, h;\n}" : "=f"(lo), "=f"(hi) : "r"(w[i]));
            o.v[2 * i] = lo; o.v[2 * i + 1] = hi;
        }
    }
    return o;
}
__device__ __forceinline__ void st8_typed(void* base, int64_t off, const f8& v, int dt) {
    if (dt == WM_DT_F32) { stg256(reinterpret_cast<float*>(base) + off, v); return; }
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (dt == WM_DT_BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(v.v[2 * i + 1]), "f"(v.v[2 * i]));
        else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(v.v[2 * i + 1]), "f"(v.v[2 * i]));
    }
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(reinterpret_cast<uint16_t*>(base) + off),
                 "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
}
inline bool dtype_ok(int dt) { return dt == WM_DT_F32 || dt == WM_DT_F16 || dt == WM_DT_BF16; }
inline size_t dtype_size(int dt) { return dt == WM_DT_F32 ? 4 : 2; }

// ---- clamp to [0,1] with torch.clamp's NaN rule (NaN propagates; fminf/fmaxf and .sat return 0 for NaN):
// two FMNMX.NAN instead of the compare + select a separate NaN test would need
__device__ __forceinline__ float clamp01_nan(float v) {
    float d;
    asm("max.NaN.f32 %0, %1, 0f00000000;\n\tmin.NaN.f32 %0, %0, 0f3F800000;" : "=f"(d) : "f"(v));
    return d;
}

// N values at once: the common case (no NaN in the group) costs N saturating moves on the FMA pipe, N-1 adds and
// ONE compare; only a group that holds a NaN (or +inf and -inf together: their sum is NaN too) takes the exact
// per-element rule.  +-inf alone saturates to 1 / 0 on the fast path, as torch.clamp does.
template <int N>
__device__ __forceinline__ void clamp01_nan_n(float* v) {
    float s = v[0];
#pragma unroll
    for (int i = 1; i < N; ++i) s += v[i];
    if (s != s) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = clamp01_nan(v[i]);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = __saturatef(v[i]);
    }
}
__device__ __forceinline__ float4 clamp01_nan4(float4 v) {
    float a[4] = {v.x, v.y, v.z, v.w};
    clamp01_nan_n<4>(a);
    return make_float4(a[0], a[1], a[2], a[3]);
}

// ---- 3-input min/max (sm_100: FMNMX3) -----------------------------------------------------
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}

// comparisons as 1.0f / 0.0f (FSET.BF): lets bit fields be accumulated with FFMA on the FMA pipe
// instead of FSETP + SEL pairs on the half-rate ALU pipe
__device__ __forceinline__ float fset_eq(float a, float b) {
    float d; asm("set.eq.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d;
}
__device__ __forceinline__ float fset_gt(float a, float b) {
    float d; asm("set.gt.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d;
}

// a / b with one Newton correction on MUFU.RCP: correctly rounded for the operand ranges
// of the quantiser (no denormals/inf) at 1/3 the issue cost of the IEEE division sequence.
__device__ __forceinline__ float div_by_recip(float a, float b, float rb) {
    float q0 = a * rb;
    float e  = fmaf(-q0, b, a);
    return fmaf(e, rb, q0);
}
__device__ __forceinline__ float fast_rcp(float b) {
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b)); return r;
}

// ---- store epilogue: out = Quantization( x + (clamp(v, 0, 1) - x) ) applied by a
// forward kernel in its own store, x read at the output position (models/IRNp_model.py:674-680).
// Passed explicitly to the forward entry points as a wm_store_epilogue (see include/wm_attack.h).
struct StoreEp { const float* x; int clamp01; int quant; int from_input; };
// from_input is set by a launcher when ep.x IS the kernel's own (dense) input: kernels that still hold the
// input value at the output position (registers / staged tile) then skip the second global read.
inline StoreEp make_store_ep(const wm_store_epilogue* e) {
    return (e && e->x) ? StoreEp{e->x, e->clamp01, e->quantize, 0} : StoreEp{nullptr, 0, 0, 0};
}
inline bool wants_store_ep(const wm_store_epilogue* e) { return e && e->x; }
#define WM_EP_CHECK(ep, who)                                                                              \
    WM_REQUIRE(!::wm::wants_store_ep(ep) || ::wm::aligned((ep)->x, 32), WM_E_ALIGN,                      \
               "%s: the store epilogue's x must be 32-byte aligned", who)
#define WM_EP_REJECT(ep, who)                                                                             \
    WM_REQUIRE(!::wm::wants_store_ep(ep), WM_E_ARG,                                                      \
               "%s: a store epilogue was passed but this code path does not apply one", who)

// n / 255 for an integer-valued n: reciprocal + one Newton step is the correctly rounded quotient
// (checked exhaustively for |n| <= 70000 against IEEE division) at 3 FMAs instead of the ~12-instruction
// division sequence with its slow path.
__device__ __forceinline__ float div255(float n) {
    constexpr float r = 1.f / 255.f;
    const float q0 = n * r;
    return fmaf(fmaf(-q0, 255.f, n), r, q0);
}
// round(v * 255) / 255 for N values at once, bit-identical to torch's (v * 255.).round() / 255.: one range
// check (and branch) for the whole group.  Fast path: the 1.5 * 2^23 trick for round-half-even (exact for
// |t| < 2^22) + div255; anything out of range takes rintf + the IEEE division.
// (the out-of-range path is a real call: inlined, its IEEE division sequence would be replicated at every store
// site of every kernel and push the stencil kernels out of the instruction cache)
static __device__ __noinline__ float quant255_slow(float t) { return __fdiv_rn(rintf(t), 255.f); }
template <int N>
__device__ __forceinline__ void quant255_n(float* v) {
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) { v[i] = __fmul_rn(v[i], 255.f); m = fmaxf(m, fabsf(v[i])); }
    if (m < 65536.f) {                   // (a NaN propagates through either path)
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = div255(__fsub_rn(__fadd_rn(v[i], 12582912.f), 12582912.f));
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = quant255_slow(v[i]);
    }
}
template <int N>
__device__ __forceinline__ void ep_apply_n(float* v, const float* x, const StoreEp& e) {
    if (e.clamp01) clamp01_nan_n<N>(v);
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = __fadd_rn(x[i], __fsub_rn(v[i], x[i]));   // same fp32 operation order as the reference
    if (e.quant) quant255_n<N>(v);
}
__device__ __forceinline__ float ep_apply(float v, float x, const StoreEp& e) {
    ep_apply_n<1>(&v, &x, e);
    return v;
}
__device__ __forceinline__ float4 ep_apply4v(float4 v, float4 x, const StoreEp& e) {
    float a[4] = {v.x, v.y, v.z, v.w};
    const float b[4] = {x.x, x.y, x.z, x.w};
    ep_apply_n<4>(a, b, e);
    return make_float4(a[0], a[1], a[2], a[3]);
}
__device__ __forceinline__ float4 ep_apply4(float4 v, const float* xp, const StoreEp& e) {
    return ep_apply4v(v, *reinterpret_cast<const float4*>(xp), e);
}

inline int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace wm
