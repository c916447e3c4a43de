// Elementwise attacks: Gaussian / GN noise, salt-and-pepper, both Dropout flavours, Cropout,
// 8-bit Quantization.  All are single-pass float4 streaming kernels; per-element randomness is
// Philox4x32-10 generated in registers (the reference draws full-size random tensors on the
// HOST and copies them over PCIe: noise_layers/salt_pepper_noise.py:14, crop.py:145,
// gaussian_noise.py:14, dropout.py:21), with an `inject` pointer so that parity tests can feed
// the very tensor the reference drew.
#include "philox.cuh"
#include "wm_common.cuh"

namespace wm {

__global__ void rng_reserve_kernel(uint64_t* state, uint64_t* slot, uint64_t count) {
    slot[0] = state[0]; slot[1] = state[1];
    state[1] += count;
}

__device__ __forceinline__ float4 ld4(const float* p, int64_t i, int64_t n) {
    if (i + 3 < n) return *reinterpret_cast<const float4*>(p + i);
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) r.x = p[i];
    if (i + 1 < n) r.y = p[i + 1];
    if (i + 2 < n) r.z = p[i + 2];
    return r;
}
__device__ __forceinline__ void st4(float* p, int64_t i, int64_t n, float4 v) {
    if (i + 3 < n) { *reinterpret_cast<float4*>(p + i) = v; return; }
    if (i < n) p[i] = v.x;
    if (i + 1 < n) p[i + 1] = v.y;
    if (i + 2 < n) p[i + 2] = v.z;
}

// typed variants (float16 / bfloat16 boundary tensors): 4 elements = one 64-bit access; tails element by element
__device__ __forceinline__ float half_bits_to_float(uint32_t h16, int dt) {
    if (dt == WM_DT_BF16) return __uint_as_float(h16 << 16);
    float f; asm("{\n .reg .b16 t;\n cvt.u16.u32 t, %1;\n cvt.f32.f16 %0, t;\n}" : "=f"(f) : "r"(h16)); return f;
}
__device__ __forceinline__ uint32_t float_to_half_bits(float v, int dt) {
    uint32_t r;
    if (dt == WM_DT_BF16) asm("{\n .reg .b16 t;\n cvt.rn.bf16.f32 t, %1;\n cvt.u32.u16 %0, t;\n}" : "=r"(r) : "f"(v));
    else asm("{\n .reg .b16 t;\n cvt.rn.f16.f32 t, %1;\n cvt.u32.u16 %0, t;\n}" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float4 ld4t(const void* p, int64_t i, int64_t n, int dt) {
    if (dt == WM_DT_F32) return ld4(reinterpret_cast<const float*>(p), i, n);
    const uint16_t* q = reinterpret_cast<const uint16_t*>(p);
    if (i + 3 < n) {
        const uint2 w = *reinterpret_cast<const uint2*>(q + i);
        return make_float4(half_bits_to_float(w.x & 0xffffu, dt), half_bits_to_float(w.x >> 16, dt),
                           half_bits_to_float(w.y & 0xffffu, dt), half_bits_to_float(w.y >> 16, dt));
    }
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) r.x = half_bits_to_float(q[i], dt);
    if (i + 1 < n) r.y = half_bits_to_float(q[i + 1], dt);
    if (i + 2 < n) r.z = half_bits_to_float(q[i + 2], dt);
    return r;
}
__device__ __forceinline__ void st4t(void* p, int64_t i, int64_t n, float4 v, int dt) {
    if (dt == WM_DT_F32) { st4(reinterpret_cast<float*>(p), i, n, v); return; }
    uint16_t* q = reinterpret_cast<uint16_t*>(p);
    const uint32_t a = float_to_half_bits(v.x, dt), b = float_to_half_bits(v.y, dt), c = float_to_half_bits(v.z, dt), d = float_to_half_bits(v.w, dt);
    if (i + 3 < n) { *reinterpret_cast<uint2*>(q + i) = make_uint2(a | (b << 16), c | (d << 16)); return; }
    if (i < n) q[i] = (uint16_t)a;
    if (i + 1 < n) q[i + 1] = (uint16_t)b;
    if (i + 2 < n) q[i + 2] = (uint16_t)c;
}

#define WM_EW_LOOP(i)                                                                         \
    for (int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n;             \
         i += int64_t(gridDim.x) * blockDim.x * 4)

// ---- Gaussian / GN ---------------------------------------------------------------------------
// fwd: y = [clamp01](x + mean + std * N);  bwd: gx = gy * 1[0 <= x + noise <= 1] (torch.clamp is
// inclusive) or gy when not clamped.
template <bool BWD, bool EP, bool TYPED = false>
__global__ void __launch_bounds__(256) gaussnoise_kernel(const void* __restrict__ x, int x_dt_, const float* __restrict__ gy,
                                                         float* __restrict__ out, int64_t n, float mean, float std,
                                                         int clamp, uint64_t seed, uint64_t offset,
                                                         const float* __restrict__ inject, const StoreEp ep) {
    resolve_rng(seed, offset);
    const Philox ph(seed);
    const int x_dt = TYPED ? x_dt_ : WM_DT_F32;     // float32 instantiation: the element-type switch folds away
    WM_EW_LOOP(i) {
        const float4 xv = ld4t(x, i, n, x_dt);      // requested first: the latency hides under Philox + Box-Muller
        float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (BWD) gv = ld4(gy, i, n);
        float4 nz;
        if (inject) nz = ld4(inject, i, n);
        else { nz = normal4(ph, (uint64_t)(i >> 2) + offset);
               nz.x = fmaf(nz.x, std, mean); nz.y = fmaf(nz.y, std, mean); nz.z = fmaf(nz.z, std, mean); nz.w = fmaf(nz.w, std, mean); }
        float4 v = make_float4(xv.x + nz.x, xv.y + nz.y, xv.z + nz.z, xv.w + nz.w);
        if (!BWD) {
            if (clamp) v = clamp01_nan4(v);
            if (EP) {
                const float4 ex = ep.from_input ? xv : ld4(ep.x, i, n);
                v = ep_apply4v(v, ex, ep);
            }
            st4(out, i, n, v);
        } else {
            float4 g = gv;
            if (clamp) { g.x = (v.x >= 0.f && v.x <= 1.f) ? g.x : 0.f; g.y = (v.y >= 0.f && v.y <= 1.f) ? g.y : 0.f;
                         g.z = (v.z >= 0.f && v.z <= 1.f) ? g.z : 0.f; g.w = (v.w >= 0.f && v.w <= 1.f) ? g.w : 0.f; }
            st4(out, i, n, g);
        }
    }
}

// Training pair of the clamped Gaussian layer: the forward also writes ONE BIT per value (the clamp's
// pass mask, 0 <= x + noise <= 1) and the backward is gx = bit ? gy : 0 — it reads neither x nor
// regenerates the noise (12 + 0.4 B/px instead of 24 B/px read, no Philox / Box-Muller).
// Layout: a warp handles 128 consecutive values per step; word 4*(i/128) + j holds, at bit l, the value
// i + 4*l + j (one ballot per j).
template <bool TYPED>
__global__ void __launch_bounds__(256) gaussnoise_mask_fwd_kernel(const void* __restrict__ x, int x_dt_, float* __restrict__ out,
                                                                  uint32_t* __restrict__ maskbits, int64_t n, float mean,
                                                                  float std, uint64_t seed, uint64_t offset,
                                                                  const float* __restrict__ inject) {
    resolve_rng(seed, offset);
    const Philox ph(seed);
    const int lane = threadIdx.x & 31;
    const int x_dt = TYPED ? x_dt_ : WM_DT_F32;
    // a warp takes 256 consecutive values per step as two independent 128-value halves: the kernel is bound by
    // instruction issue (Philox + Box-Muller), and the second half shares the loop and index arithmetic of the first
    for (int64_t base = (int64_t(blockIdx.x) * blockDim.x + (threadIdx.x & ~31)) * 8; base < n;
         base += int64_t(gridDim.x) * blockDim.x * 8) {               // warp-uniform trip count (ballots below)
        float4 xv[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) xv[h] = ld4t(x, base + 128 * h + 4 * lane, n, x_dt);    // requested first: the latency hides under Philox
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t i = base + 128 * h + 4 * lane;
            float4 nz;
            if (inject) nz = ld4(inject, i, n);
            else { nz = normal4(ph, (uint64_t)(i >> 2) + offset);
                   nz.x = fmaf(nz.x, std, mean); nz.y = fmaf(nz.y, std, mean); nz.z = fmaf(nz.z, std, mean); nz.w = fmaf(nz.w, std, mean); }
            const float4 v = make_float4(xv[h].x + nz.x, xv[h].y + nz.y, xv[h].z + nz.z, xv[h].w + nz.w);
            const float4 c = clamp01_nan4(v);           // NaN propagates (torch.clamp)
            st4(out, i, n, c);
            // 0 <= v <= 1  <=>  saturate(v) == v  (false for NaN, like torch.clamp's backward mask)
            const unsigned b0 = __ballot_sync(0xffffffffu, c.x == v.x), b1 = __ballot_sync(0xffffffffu, c.y == v.y);
            const unsigned b2 = __ballot_sync(0xffffffffu, c.z == v.z), b3 = __ballot_sync(0xffffffffu, c.w == v.w);
            if (lane == 0 && base + 128 * h < n) *reinterpret_cast<uint4*>(maskbits + ((base >> 7) + h) * 4) = make_uint4(b0, b1, b2, b3);
        }
    }
}
template <bool TYPED>
__global__ void __launch_bounds__(256) gaussnoise_mask_bwd_kernel(const float* __restrict__ gy, const uint32_t* __restrict__ maskbits,
                                                                  void* __restrict__ gx, int gx_dt_, int64_t n) {
    const int lane = threadIdx.x & 31;
    const int gx_dt = TYPED ? gx_dt_ : WM_DT_F32;
    for (int64_t base = (int64_t(blockIdx.x) * blockDim.x + (threadIdx.x & ~31)) * 4; base < n;
         base += int64_t(gridDim.x) * blockDim.x * 4) {
        const int64_t i = base + 4 * lane;
        const uint4 m = *reinterpret_cast<const uint4*>(maskbits + (base >> 7) * 4);     // one 16-byte broadcast per warp
        float4 g = ld4(gy, i, n);
        g.x = (m.x >> lane) & 1u ? g.x : 0.f; g.y = (m.y >> lane) & 1u ? g.y : 0.f;
        g.z = (m.z >> lane) & 1u ? g.z : 0.f; g.w = (m.w >> lane) & 1u ? g.w : 0.f;
        st4t(gx, i, n, g, gx_dt);
    }
}

// ---- SaltPepper (noise_layers/salt_pepper_noise.py:11-19) ------------------------------------
__device__ __forceinline__ float sp_one(float x, float r, float p0, float p1) {
    float o = r > p1 ? 0.f : x;
    return r < p0 ? 1.f : o;
}
template <bool BWD>
__global__ void __launch_bounds__(256) saltpepper_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n,
                                                         float p0, float p1, uint64_t seed, uint64_t offset,
                                                         const float* __restrict__ inject) {
    resolve_rng(seed, offset);
    const Philox ph(seed);
    WM_EW_LOOP(i) {
        const float4 r = inject ? ld4(inject, i, n) : uniform4(ph, (uint64_t)(i >> 2) + offset);
        const float4 v = ld4(x, i, n);   // x (fwd) or gy (bwd)
        float4 o;
        if (!BWD) { o.x = sp_one(v.x, r.x, p0, p1); o.y = sp_one(v.y, r.y, p0, p1);
                    o.z = sp_one(v.z, r.z, p0, p1); o.w = sp_one(v.w, r.w, p0, p1); }
        else { o.x = (r.x > p1 || r.x < p0) ? 0.f : v.x; o.y = (r.y > p1 || r.y < p0) ? 0.f : v.y;
               o.z = (r.z > p1 || r.z < p0) ? 0.f : v.z; o.w = (r.w > p1 || r.w < p0) ? 0.f : v.w; }
        st4(out, i, n, o);
    }
}

// ---- crop.Dropout (noise_layers/crop.py:142-147): y = rdn > prob ? cover : image ---------------
template <bool BWD>
__global__ void __launch_bounds__(256) dropout_elem_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                           float* __restrict__ o1, float* __restrict__ o2, int64_t n,
                                                           float prob, uint64_t seed, uint64_t offset,
                                                           const float* __restrict__ inject) {
    resolve_rng(seed, offset);
    const Philox ph(seed);
    WM_EW_LOOP(i) {
        const float4 r = inject ? ld4(inject, i, n) : uniform4(ph, (uint64_t)(i >> 2) + offset);
        if (!BWD) {
            const float4 im = ld4(a, i, n), cv = ld4(b, i, n);
            st4(o1, i, n, make_float4(r.x > prob ? cv.x : im.x, r.y > prob ? cv.y : im.y,
                                      r.z > prob ? cv.z : im.z, r.w > prob ? cv.w : im.w));
        } else {
            const float4 g = ld4(a, i, n);
            if (o1) st4(o1, i, n, make_float4(r.x > prob ? 0.f : g.x, r.y > prob ? 0.f : g.y,
                                              r.z > prob ? 0.f : g.z, r.w > prob ? 0.f : g.w));
            if (o2) st4(o2, i, n, make_float4(r.x > prob ? g.x : 0.f, r.y > prob ? g.y : 0.f,
                                              r.z > prob ? g.z : 0.f, r.w > prob ? g.w : 0.f));
        }
    }
}

// ---- dropout.Dropout (noise_layers/dropout.py:14-27): y = noised*m + cover*(1-m), m[H*W] -------
template <bool BWD>
__global__ void __launch_bounds__(256) dropout_mask_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                           const float* __restrict__ mask, float* __restrict__ o1,
                                                           float* __restrict__ o2, int64_t hw) {
    const int64_t base = int64_t(blockIdx.y) * hw;
    for (int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < hw; i += int64_t(gridDim.x) * blockDim.x * 4) {
        const float4 m = ld4(mask, i, hw);
        if (!BWD) {
            const float4 nz = ld4(a + base, i, hw), cv = ld4(b + base, i, hw);
            st4(o1 + base, i, hw, make_float4(nz.x * m.x + cv.x * (1.f - m.x), nz.y * m.y + cv.y * (1.f - m.y),
                                              nz.z * m.z + cv.z * (1.f - m.z), nz.w * m.w + cv.w * (1.f - m.w)));
        } else {
            const float4 g = ld4(a + base, i, hw);
            if (o1) st4(o1 + base, i, hw, make_float4(g.x * m.x, g.y * m.y, g.z * m.z, g.w * m.w));
            if (o2) st4(o2 + base, i, hw, make_float4(g.x * (1.f - m.x), g.y * (1.f - m.y), g.z * (1.f - m.z), g.w * (1.f - m.w)));
        }
    }
}

// H*W not a multiple of 4: planes do not start on 16-byte boundaries, one element per thread
template <bool BWD>
__global__ void __launch_bounds__(256) dropout_mask_scalar_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                  const float* __restrict__ mask, float* __restrict__ o1,
                                                                  float* __restrict__ o2, int64_t n, int64_t hw) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const float m = mask[i % hw];
        if (!BWD) o1[i] = a[i] * m + b[i] * (1.f - m);           // same expression as the vector kernel
        else { if (o1) o1[i] = a[i] * m; if (o2) o2[i] = a[i] * (1.f - m); }
    }
}

__global__ void __launch_bounds__(256) bernoulli_kernel(float* __restrict__ mask, int64_t n, float keep,
                                                        uint64_t seed, uint64_t offset) {
    resolve_rng(seed, offset);
    const Philox ph(seed);
    WM_EW_LOOP(i) {
        const float4 r = uniform4(ph, (uint64_t)(i >> 2) + offset);
        st4(mask, i, n, make_float4(r.x < keep ? 1.f : 0.f, r.y < keep ? 1.f : 0.f, r.z < keep ? 1.f : 0.f, r.w < keep ? 1.f : 0.f));
    }
}

// ---- Quantization (models/modules/Quantization.py:7-10; utils/JPEG_utils.py:46-50 with clamp) --
__global__ void __launch_bounds__(256) quantize8_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, int clamp01) {
    WM_EW_LOOP(i) {
        float4 v = ld4(x, i, n);
        if (clamp01) v = clamp01_nan4(v);
        // correctly rounded quotient: bit-parity with torch's `/ 255.` (quant255_n, wm_common.cuh)
        float q[4] = {v.x, v.y, v.z, v.w};
        quant255_n<4>(q);
        st4(y, i, n, make_float4(q[0], q[1], q[2], q[3]));
    }
}

// ---- Cropout (noise_layers/crop.py:128-134) ----------------------------------------------------
__global__ void __launch_bounds__(256) cropout_kernel(const float* __restrict__ image, const float* __restrict__ cover,
                                                      float* __restrict__ y, int64_t total, int H, int W,
                                                      int h0, int h1, int w0, int w1) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int w = int(i % W), h = int((i / W) % H);
        const bool in = h >= h0 && h < h1 && w >= w0 && w < w1;
        y[i] = in ? image[i] : cover[i];
    }
}

// ---- post-attack epilogue (models/IRNp_model.py:674-680) --------------------------------------
//   sim = clamp(attack(x), 0, 1);  attacked = x + (sim - x).detach();  out = Quantization(attacked)
// = four elementwise torch passes + a torch.cat in the reference; here one pass that can write
// straight into a slice of the K-way batch.  Same fp32 operation order (no FMA contraction), so the
// values are bit-identical to the reference's.  Backward is the identity on x (straight-through).
__device__ __forceinline__ float epilogue1(float x, float s, int clamp01, int quant) {
    if (clamp01) s = clamp01_nan(s);
    float v = __fadd_rn(x, __fsub_rn(s, x));
    if (quant) v = __fdiv_rn(rintf(__fmul_rn(v, 255.f)), 255.f);
    return v;
}
__global__ void __launch_bounds__(256) attack_epilogue_kernel(const float* __restrict__ x, const float* __restrict__ sim,
                                                              float* __restrict__ out, int64_t n, int clamp01, int quant) {
    WM_EW_LOOP(i) {
        const float4 a = ld4(x, i, n), b = ld4(sim, i, n);
        float q[4] = {epilogue1(a.x, b.x, clamp01, 0), epilogue1(a.y, b.y, clamp01, 0),
                      epilogue1(a.z, b.z, clamp01, 0), epilogue1(a.w, b.w, clamp01, 0)};
        if (quant) quant255_n<4>(q);
        st4(out, i, n, make_float4(q[0], q[1], q[2], q[3]));
    }
}
// slices of a K-way batch whose per-image element count is odd do not start on 16-byte boundaries: one element per thread
__global__ void __launch_bounds__(256) attack_epilogue_scalar_kernel(const float* __restrict__ x, const float* __restrict__ sim,
                                                                     float* __restrict__ out, int64_t n, int clamp01, int quant) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        float q[1] = {epilogue1(x[i], sim[i], clamp01, 0)};
        if (quant) quant255_n<1>(q);
        out[i] = q[0];
    }
}
__global__ void __launch_bounds__(256) slice_sum_scalar_kernel(const float* __restrict__ g, float* __restrict__ out, int64_t n, int K) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        float acc = g[i];
        for (int k = 1; k < K; ++k) acc += g[int64_t(k) * n + i];
        out[i] = acc;
    }
}
// out[i] = sum_k g[k * n + i]   (straight-through backward of the K-way bank: every slice passes gy to x)
__global__ void __launch_bounds__(256) slice_sum_kernel(const float* __restrict__ g, float* __restrict__ out, int64_t n, int K) {
    WM_EW_LOOP(i) {
        float4 acc = ld4(g, i, n);
        for (int k = 1; k < K; ++k) {
            const float4 v = ld4(g + int64_t(k) * n, i, n);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        st4(out, i, n, acc);
    }
}

// ---- hybrid attack: convex mix of K attacked versions (models/IRNcrop_model.py:357-373, as intended) ------------
//   mixed = sum_k alpha[b, k] * y_k;  out = Quantization( mixed + (clamp(mixed, 0, 1) - mixed).detach() )
// = K multiplies + K - 1 adds + the clamp_with_grad / Quantization passes of the trainer (~430 B/px at K = 5); here the
// K tensors are read once and one is written (12 K + 12 B/px).  Same fp32 operation order (products added left to
// right, no FMA contraction): bit-identical.  Backward: g_k = alpha[b, k] * gy (clamp_with_grad and Quantization are
// straight-through).
struct MixArgs { const float* y[WM_MIX_MAX]; float* g[WM_MIX_MAX]; const float* alpha; int K, clamp01, quant; int64_t chw, n; };

__device__ __forceinline__ float mix_finish(float v, int clamp01) {
    return clamp01 ? __fadd_rn(v, __fsub_rn(clamp01_nan(v), v)) : v;
}
template <bool VEC>
__global__ void __launch_bounds__(256) mix_fwd_kernel(const MixArgs a, float* __restrict__ out) {
    constexpr int E = VEC ? 4 : 1;
    for (int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * E; i < a.n; i += int64_t(gridDim.x) * blockDim.x * E) {
        const int64_t b = i / a.chw;
        float acc[E];
#pragma unroll
        for (int k = 0; k < WM_MIX_MAX; ++k) {
            if (k < a.K) {
                const float w = __ldg(a.alpha + b * a.K + k);
                float v[E];
                if (VEC) { const float4 t = *reinterpret_cast<const float4*>(a.y[k] + i); v[0] = t.x; v[1 % E] = t.y; v[2 % E] = t.z; v[3 % E] = t.w; }
                else v[0] = a.y[k][i];
#pragma unroll
                for (int e = 0; e < E; ++e) acc[e] = k == 0 ? __fmul_rn(w, v[e]) : __fadd_rn(acc[e], __fmul_rn(w, v[e]));
            }
        }
#pragma unroll
        for (int e = 0; e < E; ++e) acc[e] = mix_finish(acc[e], a.clamp01);
        if (a.quant) quant255_n<E>(acc);
        if (VEC) *reinterpret_cast<float4*>(out + i) = make_float4(acc[0], acc[1 % E], acc[2 % E], acc[3 % E]);
        else out[i] = acc[0];
    }
}
template <bool VEC>
__global__ void __launch_bounds__(256) mix_bwd_kernel(const MixArgs a, const float* __restrict__ gy) {
    constexpr int E = VEC ? 4 : 1;
    for (int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * E; i < a.n; i += int64_t(gridDim.x) * blockDim.x * E) {
        const int64_t b = i / a.chw;
        float g[E];
        if (VEC) { const float4 t = *reinterpret_cast<const float4*>(gy + i); g[0] = t.x; g[1 % E] = t.y; g[2 % E] = t.z; g[3 % E] = t.w; }
        else g[0] = gy[i];
#pragma unroll
        for (int k = 0; k < WM_MIX_MAX; ++k) {
            if (k < a.K && a.g[k]) {
                const float w = __ldg(a.alpha + b * a.K + k);
                if (VEC) *reinterpret_cast<float4*>(a.g[k] + i) = make_float4(__fmul_rn(w, g[0]), __fmul_rn(w, g[1 % E]), __fmul_rn(w, g[2 % E]), __fmul_rn(w, g[3 % E]));
                else a.g[k][i] = __fmul_rn(w, g[0]);
            }
        }
    }
}

// ---- tamper / splice (models/IRNcrop_model.py:348, models/IRNp_model.py:600) ---------------------
//   out = a * (1 - m) + b * m,  m: [B, 1, H, W] broadcast over the C channels of a, b: [B, C, H, W]
// bwd: ga = gy * (1 - m), gb = gy * m.  hw % 4 == 0 keeps a float4 inside one plane.
template <bool BWD>
__global__ void __launch_bounds__(256) splice_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                     const float* __restrict__ m, float* __restrict__ o1, float* __restrict__ o2,
                                                     int64_t n, int64_t hw, int C) {
    WM_EW_LOOP(i) {
        const int64_t plane = i / hw, bi = plane / C;
        const float4 mm = *reinterpret_cast<const float4*>(m + bi * hw + (i - plane * hw));
        const float4 va = *reinterpret_cast<const float4*>(a + i);
        const float mv[4] = {mm.x, mm.y, mm.z, mm.w}, av[4] = {va.x, va.y, va.z, va.w};
        float r1[4], r2[4];
        if (!BWD) {
            const float4 vb = *reinterpret_cast<const float4*>(b + i);
            const float bv[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                r1[k] = __fadd_rn(__fmul_rn(av[k], __fsub_rn(1.f, mv[k])), __fmul_rn(bv[k], mv[k]));
            *reinterpret_cast<float4*>(o1 + i) = make_float4(r1[0], r1[1], r1[2], r1[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) { r1[k] = __fmul_rn(av[k], __fsub_rn(1.f, mv[k])); r2[k] = __fmul_rn(av[k], mv[k]); }
            if (o1) *reinterpret_cast<float4*>(o1 + i) = make_float4(r1[0], r1[1], r1[2], r1[3]);
            if (o2) *reinterpret_cast<float4*>(o2 + i) = make_float4(r2[0], r2[1], r2[2], r2[3]);
        }
    }
}

template <bool BWD>
__global__ void __launch_bounds__(256) splice_scalar_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            const float* __restrict__ m, float* __restrict__ o1, float* __restrict__ o2,
                                                            int64_t n, int64_t hw, int C) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t plane = i / hw, bi = plane / C;
        const float mv = m[bi * hw + (i - plane * hw)], av = a[i];
        if (!BWD) o1[i] = __fadd_rn(__fmul_rn(av, __fsub_rn(1.f, mv)), __fmul_rn(b[i], mv));
        else { if (o1) o1[i] = __fmul_rn(av, __fsub_rn(1.f, mv)); if (o2) o2[i] = __fmul_rn(av, mv); }
    }
}

// ---- uint8 frames -> [0,1] float (the data format on the host side of the path) -----------------
// The reference's loaders divide 8-bit frames by 255 on the CPU and upload fp32 (data/Dataloader.py);
// uploading the bytes and converting on the device moves 4x less over PCIe.  Same value as
// torch's  u8.float() / 255  (true division).
__global__ void __launch_bounds__(256) u8_to_unit_float_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 16; i < n; i += int64_t(gridDim.x) * blockDim.x * 16) {
        if (i + 15 < n) {
            const uint4 w = *reinterpret_cast<const uint4*>(src + i);
            const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                *reinterpret_cast<float4*>(dst + i + 4 * k) =
                    // div255: reciprocal + one Newton step == the IEEE quotient for these integers (wm_common.cuh)
                    make_float4(div255(float(ws[k] & 0xff)), div255(float((ws[k] >> 8) & 0xff)),
                                div255(float((ws[k] >> 16) & 0xff)), div255(float(ws[k] >> 24)));
        } else {
            for (int64_t j = i; j < n; ++j) dst[j] = __fdiv_rn(float(src[j]), 255.f);
        }
    }
}

// The opposite direction: attacked frames leave the device as bytes (what a video encoder / an 8-bit
// evaluation pipeline consumes): dst = rint(clamp(src, 0, 1) * 255), round-half-even like torch.round,
// i.e. the integer k of the Quantization layer's value k/255 (models/modules/Quantization.py:9); NaN -> 0.
__global__ void __launch_bounds__(256) unit_float_to_u8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 16; i < n; i += int64_t(gridDim.x) * blockDim.x * 16) {
        if (i + 15 < n) {
            uint32_t ws[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 v = ldg128_stream(src + i + 4 * k);
                const float f[4] = {v.x, v.y, v.z, v.w};
                uint32_t w = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    w |= uint32_t(__float2int_rn(__fmul_rn(__saturatef(f[j]), 255.f))) << (8 * j);
                ws[k] = w;
            }
            *reinterpret_cast<uint4*>(dst + i) = make_uint4(ws[0], ws[1], ws[2], ws[3]);
        } else {
            for (int64_t j = i; j < n; ++j) dst[j] = uint8_t(__float2int_rn(__fmul_rn(__saturatef(src[j]), 255.f)));
        }
    }
}

static inline unsigned ew_grid(int64_t n_vec) {
    const int64_t want = (n_vec + 255) / 256;
    const int64_t cap = int64_t(sm_count()) * 16;
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace wm

using namespace wm;

#define EW_ALIGN_CHECK(who, ...)                                                                \
    do { const void* ps__[] = {__VA_ARGS__};                                                    \
         for (const void* p__ : ps__) WM_REQUIRE(p__ == nullptr || aligned(p__, 16), WM_E_ALIGN, \
             "%s: pointers must be 16-byte aligned", who); } while (0)

extern "C" int wm_gaussnoise_fwd(const void* x, int x_dtype, float* y, int64_t n, float mean, float std, int clamp,
                                 uint64_t seed, uint64_t offset, const float* inject,
                                 const wm_store_epilogue* ep_in, void* stream) {
    if (n <= 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(x && y, WM_E_NULL, "wm_gaussnoise_fwd: null pointer");
    WM_REQUIRE(dtype_ok(x_dtype), WM_E_ARG, "wm_gaussnoise_fwd: unknown element type %d", x_dtype);
    WM_EP_CHECK(ep_in, "wm_gaussnoise_fwd");
    EW_ALIGN_CHECK("wm_gaussnoise_fwd", x, y, inject);
    if (n <= 0) return WM_OK;
    StoreEp ep = make_store_ep(ep_in);
    ep.from_input = ep.x == x && x_dtype == WM_DT_F32;
    WM_REQUIRE(!ep.x || x_dtype == WM_DT_F32, WM_E_ARG, "wm_gaussnoise_fwd: the store epilogue needs a float32 image");
    if (x_dtype != WM_DT_F32) gaussnoise_kernel<false, false, true><<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(x, x_dtype, nullptr, y, n, mean, std, clamp, seed, offset, inject, ep);
    else if (ep.x) gaussnoise_kernel<false, true><<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(x, x_dtype, nullptr, y, n, mean, std, clamp, seed, offset, inject, ep);
    else gaussnoise_kernel<false, false><<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(x, x_dtype, nullptr, y, n, mean, std, clamp, seed, offset, inject, ep);
    WM_LAUNCH_CHECK("wm_gaussnoise_fwd");
    return WM_OK;
}
extern "C" int wm_gaussnoise_bwd(const float* x, const float* gy, float* gx, int64_t n, float mean, float std,
                                 int clamp, uint64_t seed, uint64_t offset, const float* inject, void* stream) {
    if (n <= 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(gy && gx && (x || !clamp), WM_E_NULL, "wm_gaussnoise_bwd: null pointer");
    EW_ALIGN_CHECK("wm_gaussnoise_bwd", x, gy, gx, inject);
    if (n <= 0) return WM_OK;
    if (!clamp) {   // GN: identity gradient
        cudaError_t e = cudaMemcpyAsync(gx, gy, sizeof(float) * n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
        return e == cudaSuccess ? WM_OK : cuda_fail(e, "wm_gaussnoise_bwd");
    }
    gaussnoise_kernel<true, false><<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(x, WM_DT_F32, gy, gx, n, mean, std, clamp, seed, offset, inject,
                                                                                  StoreEp{nullptr, 0, 0});
    WM_LAUNCH_CHECK("wm_gaussnoise_bwd");
    return WM_OK;
}
extern "C" int wm_gaussnoise_fwd_mask(const void* x, int x_dtype, float* y, uint32_t* maskbits, int64_t n, float mean, float std,
                                      uint64_t seed, uint64_t offset, const float* inject, void* stream) {
    if (n <= 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(x && y && maskbits, WM_E_NULL, "wm_gaussnoise_fwd_mask: null pointer");
    WM_REQUIRE(dtype_ok(x_dtype), WM_E_ARG, "wm_gaussnoise_fwd_mask: unknown element type %d", x_dtype);
    EW_ALIGN_CHECK("wm_gaussnoise_fwd_mask", x, y, inject, maskbits);
    if (x_dtype != WM_DT_F32) gaussnoise_mask_fwd_kernel<true><<<ew_grid((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, x_dtype, y, maskbits, n, mean, std, seed, offset, inject);
    else gaussnoise_mask_fwd_kernel<false><<<ew_grid((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, x_dtype, y, maskbits, n, mean, std, seed, offset, inject);
    WM_LAUNCH_CHECK("wm_gaussnoise_fwd_mask");
    return WM_OK;
}
extern "C" int wm_gaussnoise_bwd_mask(const float* gy, const uint32_t* maskbits, void* gx, int gx_dtype, int64_t n, void* stream) {
    if (n <= 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(gy && gx && maskbits, WM_E_NULL, "wm_gaussnoise_bwd_mask: null pointer");
    WM_REQUIRE(dtype_ok(gx_dtype), WM_E_ARG, "wm_gaussnoise_bwd_mask: unknown element type %d", gx_dtype);
    EW_ALIGN_CHECK("wm_gaussnoise_bwd_mask", gy, gx, maskbits);
    if (gx_dtype != WM_DT_F32) gaussnoise_mask_bwd_kernel<true><<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(gy, maskbits, gx, gx_dtype, n);
    else gaussnoise_mask_bwd_kernel<false><<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(gy, maskbits, gx, gx_dtype, n);
    WM_LAUNCH_CHECK("wm_gaussnoise_bwd_mask");
    return WM_OK;
}
extern "C" int wm_rng_reserve(uint64_t* state, uint64_t* slot, uint64_t count, void* stream) {
    WM_REQUIRE(state && slot, WM_E_NULL, "wm_rng_reserve: null pointer");
    WM_REQUIRE(aligned(state, 8) && aligned(slot, 8), WM_E_ALIGN, "wm_rng_reserve: pointers must be 8-byte aligned");
    rng_reserve_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state, slot, count);
    WM_LAUNCH_CHECK("wm_rng_reserve");
    return WM_OK;
}
extern "C" int wm_saltpepper_fwd(const float* x, float* y, int64_t n, float prob, uint64_t seed, uint64_t offset,
                                 const float* inject, void* stream) {
    if (n <= 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(x && y, WM_E_NULL, "wm_saltpepper_fwd: null pointer");
    EW_ALIGN_CHECK("wm_saltpepper_fwd", x, y, inject);
    if (n <= 0) return WM_OK;
    // thresholds in the reference's arithmetic: python doubles prob/2 and 1 - prob/2 compared
    // against a float32 tensor (promoted scalar -> float32)
    const float p0 = (float)((double)prob / 2.0), p1 = (float)(1.0 - (double)prob / 2.0);
    saltpepper_kernel<false><<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(x, y, n, p0, p1, seed, offset, inject);
    WM_LAUNCH_CHECK("wm_saltpepper_fwd");
    return WM_OK;
}
extern "C" int wm_saltpepper_bwd(const float* gy, float* gx, int64_t n, float prob, uint64_t seed, uint64_t offset,
                                 const float* inject, void* stream) {
    if (n <= 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(gy && gx, WM_E_NULL, "wm_saltpepper_bwd: null pointer");
    EW_ALIGN_CHECK("wm_saltpepper_bwd", gy, gx, inject);
    if (n <= 0) return WM_OK;
    const float p0 = (float)((double)prob / 2.0), p1 = (float)(1.0 - (double)prob / 2.0);
    saltpepper_kernel<true><<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(gy, gx, n, p0, p1, seed, offset, inject);
    WM_LAUNCH_CHECK("wm_saltpepper_bwd");
    return WM_OK;
}
extern "C" int wm_dropout_elem_fwd(const float* image, const float* cover, float* y, int64_t n, float prob,
                                   uint64_t seed, uint64_t offset, const float* inject, void* stream) {
    if (n <= 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(image && cover && y, WM_E_NULL, "wm_dropout_elem_fwd: null pointer");
    EW_ALIGN_CHECK("wm_dropout_elem_fwd", image, cover, y, inject);
    if (n <= 0) return WM_OK;
    dropout_elem_kernel<false><<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(image, cover, y, nullptr, n, prob, seed, offset, inject);
    WM_LAUNCH_CHECK("wm_dropout_elem_fwd");
    return WM_OK;
}
extern "C" int wm_dropout_elem_bwd(const float* gy, float* g_image, float* g_cover, int64_t n, float prob,
                                   uint64_t seed, uint64_t offset, const float* inject, void* stream) {
    if (n <= 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(gy && (g_image || g_cover), WM_E_NULL, "wm_dropout_elem_bwd: null pointer");
    EW_ALIGN_CHECK("wm_dropout_elem_bwd", gy, g_image, g_cover, inject);
    if (n <= 0) return WM_OK;
    dropout_elem_kernel<true><<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(gy, nullptr, g_image, g_cover, n, prob, seed, offset, inject);
    WM_LAUNCH_CHECK("wm_dropout_elem_bwd");
    return WM_OK;
}
extern "C" int wm_dropout_mask_fwd(const float* noised, const float* cover, const float* mask_hw, float* y,
                                   int64_t planes, int64_t hw, void* stream) {
    if (planes * hw <= 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(noised && cover && mask_hw && y, WM_E_NULL, "wm_dropout_mask_fwd: null pointer");
    EW_ALIGN_CHECK("wm_dropout_mask_fwd", noised, cover, mask_hw, y);
    if (planes <= 0 || hw <= 0) return WM_OK;
    if (hw % 4 != 0 && planes != 1) {
        dropout_mask_scalar_kernel<false><<<ew_grid(planes * hw), 256, 0, (cudaStream_t)stream>>>(noised, cover, mask_hw, y, nullptr, planes * hw, hw);
        WM_LAUNCH_CHECK("wm_dropout_mask_fwd");
        return WM_OK;
    }
    dim3 grid(ew_grid((hw + 3) / 4), (unsigned)planes);
    dropout_mask_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(noised, cover, mask_hw, y, nullptr, hw);
    WM_LAUNCH_CHECK("wm_dropout_mask_fwd");
    return WM_OK;
}
extern "C" int wm_dropout_mask_bwd(const float* gy, const float* mask_hw, float* g_noised, float* g_cover,
                                   int64_t planes, int64_t hw, void* stream) {
    if (planes * hw <= 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(gy && mask_hw && (g_noised || g_cover), WM_E_NULL, "wm_dropout_mask_bwd: null pointer");
    EW_ALIGN_CHECK("wm_dropout_mask_bwd", gy, mask_hw, g_noised, g_cover);
    if (planes <= 0 || hw <= 0) return WM_OK;
    if (hw % 4 != 0 && planes != 1) {
        dropout_mask_scalar_kernel<true><<<ew_grid(planes * hw), 256, 0, (cudaStream_t)stream>>>(gy, nullptr, mask_hw, g_noised, g_cover, planes * hw, hw);
        WM_LAUNCH_CHECK("wm_dropout_mask_bwd");
        return WM_OK;
    }
    dim3 grid(ew_grid((hw + 3) / 4), (unsigned)planes);
    dropout_mask_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(gy, nullptr, mask_hw, g_noised, g_cover, hw);
    WM_LAUNCH_CHECK("wm_dropout_mask_bwd");
    return WM_OK;
}
extern "C" int wm_bernoulli_mask(float* mask_hw, int64_t hw, float keep, uint64_t seed, uint64_t offset, void* stream) {
    if (hw <= 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(mask_hw, WM_E_NULL, "wm_bernoulli_mask: null pointer");
    EW_ALIGN_CHECK("wm_bernoulli_mask", mask_hw);
    if (hw <= 0) return WM_OK;
    bernoulli_kernel<<<ew_grid((hw + 3) / 4), 256, 0, (cudaStream_t)stream>>>(mask_hw, hw, keep, seed, offset);
    WM_LAUNCH_CHECK("wm_bernoulli_mask");
    return WM_OK;
}
extern "C" int wm_quantize8_fwd(const float* x, float* y, int64_t n, int clamp01, void* stream) {
    if (n <= 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(x && y, WM_E_NULL, "wm_quantize8_fwd: null pointer");
    EW_ALIGN_CHECK("wm_quantize8_fwd", x, y);
    if (n <= 0) return WM_OK;
    quantize8_kernel<<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(x, y, n, clamp01);
    WM_LAUNCH_CHECK("wm_quantize8_fwd");
    return WM_OK;
}
extern "C" int wm_cropout_fwd(const float* image, const float* cover, float* y, int64_t planes, int H, int W,
                              int h0, int h1, int w0, int w1, void* stream) {
    if (planes * H * W <= 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(image && cover && y, WM_E_NULL, "wm_cropout_fwd: null pointer");
    const int64_t total = planes * H * W;
    if (total <= 0) return WM_OK;
    cropout_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(image, cover, y, total, H, W, h0, h1, w0, w1);
    WM_LAUNCH_CHECK("wm_cropout_fwd");
    return WM_OK;
}

extern "C" int wm_attack_epilogue_fwd(const float* x, const float* sim, float* out, int64_t n, int clamp01, int quantize,
                                      void* stream) {
    if (n <= 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(x && sim && out, WM_E_NULL, "wm_attack_epilogue_fwd: null pointer");
    if (n <= 0) return WM_OK;
    if (!(aligned(x, 16) && aligned(sim, 16) && aligned(out, 16))) {
        WM_REQUIRE(aligned(x, 4) && aligned(sim, 4) && aligned(out, 4), WM_E_ALIGN, "wm_attack_epilogue_fwd: pointers must be 4-byte aligned");
        attack_epilogue_scalar_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, sim, out, n, clamp01, quantize);
        WM_LAUNCH_CHECK("wm_attack_epilogue_fwd");
        return WM_OK;
    }
    attack_epilogue_kernel<<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(x, sim, out, n, clamp01, quantize);
    WM_LAUNCH_CHECK("wm_attack_epilogue_fwd");
    return WM_OK;
}
extern "C" int wm_slice_sum(const float* g, float* out, int64_t n, int K, void* stream) {
    if (n <= 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(g && out, WM_E_NULL, "wm_slice_sum: null pointer");
    WM_REQUIRE(K >= 1, WM_E_ARG, "wm_slice_sum: K >= 1 required (K=%d)", K);
    if (n <= 0) return WM_OK;
    if (n % 4 != 0 || !aligned(g, 16) || !aligned(out, 16)) {       // odd slices: not on 16-byte boundaries
        slice_sum_scalar_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(g, out, n, K);
        WM_LAUNCH_CHECK("wm_slice_sum");
        return WM_OK;
    }
    slice_sum_kernel<<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(g, out, n, K);
    WM_LAUNCH_CHECK("wm_slice_sum");
    return WM_OK;
}
static int mix_fill(MixArgs& a, const wm_mix_desc* d, const float* alpha, int64_t B, int64_t chw, bool bwd, const char* who) {
    WM_REQUIRE(d && alpha, WM_E_NULL, "%s: null pointer", who);
    WM_REQUIRE(d->K >= 1 && d->K <= WM_MIX_MAX, WM_E_ARG, "%s: K must be 1..%d (got %d)", who, WM_MIX_MAX, d->K);
    a.alpha = alpha; a.K = d->K; a.clamp01 = d->clamp01; a.quant = d->quantize; a.chw = chw; a.n = B * chw;
    for (int k = 0; k < WM_MIX_MAX; ++k) { a.y[k] = nullptr; a.g[k] = nullptr; }
    for (int k = 0; k < d->K; ++k) {
        WM_REQUIRE(bwd || d->t[k], WM_E_NULL, "%s: member %d is null", who, k);
        WM_REQUIRE(aligned(d->t[k], 4), WM_E_ALIGN, "%s: member %d must be 4-byte aligned", who, k);
        if (bwd) a.g[k] = d->t[k]; else a.y[k] = d->t[k];
    }
    return WM_OK;
}
static bool mix_vec_ok(const wm_mix_desc* d, const void* other, int64_t chw) {
    if (chw % 4 != 0 || !aligned(other, 16)) return false;
    for (int k = 0; k < d->K; ++k) if (!aligned(d->t[k], 16)) return false;
    return true;
}
extern "C" int wm_mix_fwd(const wm_mix_desc* desc_host, const float* alpha, float* out, int64_t B, int64_t chw, void* stream) {
    if (B * chw <= 0) return WM_OK;      // empty work: nothing to validate or launch
    MixArgs a{};
    if (int rc = mix_fill(a, desc_host, alpha, B, chw, false, "wm_mix_fwd")) return rc;
    WM_REQUIRE(out && aligned(out, 4), WM_E_NULL, "wm_mix_fwd: null / misaligned output");
    if (mix_vec_ok(desc_host, out, chw)) mix_fwd_kernel<true><<<ew_grid(a.n / 4), 256, 0, (cudaStream_t)stream>>>(a, out);
    else mix_fwd_kernel<false><<<ew_grid(a.n), 256, 0, (cudaStream_t)stream>>>(a, out);
    WM_LAUNCH_CHECK("wm_mix_fwd");
    return WM_OK;
}
extern "C" int wm_mix_bwd(const float* gy, const float* alpha, const wm_mix_desc* grads_host, int64_t B, int64_t chw, void* stream) {
    if (B * chw <= 0) return WM_OK;      // empty work: nothing to validate or launch
    MixArgs a{};
    if (int rc = mix_fill(a, grads_host, alpha, B, chw, true, "wm_mix_bwd")) return rc;
    WM_REQUIRE(gy && aligned(gy, 4), WM_E_NULL, "wm_mix_bwd: null / misaligned gy");
    if (mix_vec_ok(grads_host, gy, chw)) mix_bwd_kernel<true><<<ew_grid(a.n / 4), 256, 0, (cudaStream_t)stream>>>(a, gy);
    else mix_bwd_kernel<false><<<ew_grid(a.n), 256, 0, (cudaStream_t)stream>>>(a, gy);
    WM_LAUNCH_CHECK("wm_mix_bwd");
    return WM_OK;
}
extern "C" int wm_splice_fwd(const float* a, const float* b, const float* mask, float* out, int64_t B, int C, int64_t hw,
                             void* stream) {
    if (B * C * hw <= 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(a && b && mask && out, WM_E_NULL, "wm_splice_fwd: null pointer");
    WM_REQUIRE(C >= 1, WM_E_SHAPE, "wm_splice_fwd: C must be >= 1");
    EW_ALIGN_CHECK("wm_splice_fwd", a, b, mask, out);
    const int64_t n = B * C * hw;
    if (n <= 0) return WM_OK;
    if (hw % 4 != 0) {
        splice_scalar_kernel<false><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(a, b, mask, out, nullptr, n, hw, C);
        WM_LAUNCH_CHECK("wm_splice_fwd");
        return WM_OK;
    }
    splice_kernel<false><<<ew_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(a, b, mask, out, nullptr, n, hw, C);
    WM_LAUNCH_CHECK("wm_splice_fwd");
    return WM_OK;
}
extern "C" int wm_splice_bwd(const float* gy, const float* mask, float* ga, float* gb, int64_t B, int C, int64_t hw,
                             void* stream) {
    if (B * C * hw <= 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(gy && mask && (ga || gb), WM_E_NULL, "wm_splice_bwd: null pointer");
    WM_REQUIRE(C >= 1, WM_E_SHAPE, "wm_splice_bwd: C must be >= 1");
    EW_ALIGN_CHECK("wm_splice_bwd", gy, mask, ga, gb);
    const int64_t n = B * C * hw;
    if (n <= 0) return WM_OK;
    if (hw % 4 != 0) {
        splice_scalar_kernel<true><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(gy, nullptr, mask, ga, gb, n, hw, C);
        WM_LAUNCH_CHECK("wm_splice_bwd");
        return WM_OK;
    }
    splice_kernel<true><<<ew_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(gy, nullptr, mask, ga, gb, n, hw, C);
    WM_LAUNCH_CHECK("wm_splice_bwd");
    return WM_OK;
}

extern "C" int wm_u8_to_unit_float(const uint8_t* src, float* dst, int64_t n, void* stream) {
    if (n <= 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(src && dst, WM_E_NULL, "wm_u8_to_unit_float: null pointer");
    WM_REQUIRE(aligned(src, 16) && aligned(dst, 16), WM_E_ALIGN, "wm_u8_to_unit_float: pointers must be 16-byte aligned");
    if (n <= 0) return WM_OK;
    u8_to_unit_float_kernel<<<ew_grid((n + 15) / 16), 256, 0, (cudaStream_t)stream>>>(src, dst, n);
    WM_LAUNCH_CHECK("wm_u8_to_unit_float");
    return WM_OK;
}

extern "C" int wm_unit_float_to_u8(const float* src, uint8_t* dst, int64_t n, void* stream) {
    if (n <= 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(src && dst, WM_E_NULL, "wm_unit_float_to_u8: null pointer");
    WM_REQUIRE(aligned(src, 16) && aligned(dst, 16), WM_E_ALIGN, "wm_unit_float_to_u8: pointers must be 16-byte aligned");
    unit_float_to_u8_kernel<<<ew_grid((n + 15) / 16), 256, 0, (cudaStream_t)stream>>>(src, dst, n);
    WM_LAUNCH_CHECK("wm_unit_float_to_u8");
    return WM_OK;
}
