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

// ---- 3-input min/max (sm_100: FMNMX3) -----------------------------------------------------
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
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
