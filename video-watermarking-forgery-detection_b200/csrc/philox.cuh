// Philox4x32-10 counter-based generator + the uniform / normal draws of the stochastic attacks, shared by
// elementwise.cu (stand-alone layers) and bank3.cu (shared-read bank) so that both produce the SAME numbers
// for the same (seed, counter).
#pragma once
#include "wm_common.cuh"

namespace wm {

// ---- Philox4x32-10 (Salmon et al. 2011), counter = (idx_lo, idx_hi, 0, 0), key = seed -------
struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    __device__ __forceinline__ uint4 operator()(uint64_t ctr) const {
        uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x243F6A88u, c3 = 0x85A308D3u;
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ a; c1 = lo1; c2 = hi0 ^ c3 ^ b; c3 = lo0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

// Device-resident randomness (CUDA-graph capture): when seed == WM_RNG_FROM_DEVICE the `offset` argument is
// a device pointer to {seed, offset} — written by wm_rng_reserve in the same stream — so that a captured
// launch draws fresh numbers at every replay and its backward regenerates exactly the same ones.
__device__ __forceinline__ void resolve_rng(uint64_t& seed, uint64_t& offset) {
    if (seed == WM_RNG_FROM_DEVICE) {
        const uint64_t* p = reinterpret_cast<const uint64_t*>(offset);
        seed = __ldg(p); offset = __ldg(p + 1);
    }
}
// uniform in [0,1) with 24 bits (same support as torch.rand float32)
__device__ __forceinline__ float u01(uint32_t r) { return (r >> 8) * (1.0f / 16777216.0f); }

__device__ __forceinline__ float4 uniform4(const Philox& ph, uint64_t ctr) {
    const uint4 r = ph(ctr);
    return make_float4(u01(r.x), u01(r.y), u01(r.z), u01(r.w));
}
__device__ __forceinline__ float4 normal4(const Philox& ph, uint64_t ctr) {
    const uint4 r = ph(ctr);
    // Box-Muller on (0,1] x [0,1].  Radius uniforms (r + 0.5) * 2^-32: one unsigned conversion + one FFMA (the conversion
    // rounds, 2^32 - 1 lands on 1.0, 0 on 2^-33: radius <= 6.8 sigma); angle uniforms r * 2^-32.
    const float u1 = fmaf(__uint2float_rn(r.x), 0x1p-32f, 0x1p-33f), u2 = __uint2float_rn(r.y) * 0x1p-32f;
    const float u3 = fmaf(__uint2float_rn(r.z), 0x1p-32f, 0x1p-33f), u4 = __uint2float_rn(r.w) * 0x1p-32f;
    // -2 ln u = (-2 ln 2) log2 u: MUFU.LG2 + one FMUL; sqrt.approx (MUFU.SQRT, max 1 ulp): the radius of a RANDOM draw
    // needs no IEEE rounding
    float ra, rb;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(-1.3862943611198906f * __log2f(u1)));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(-1.3862943611198906f * __log2f(u3)));
    float s1, c1, s2, c2;
    __sincosf(6.283185307179586f * u2, &s1, &c1);
    __sincosf(6.283185307179586f * u4, &s2, &c2);
    return make_float4(ra * c1, ra * s1, rb * c2, rb * s2);
}

}  // namespace wm
