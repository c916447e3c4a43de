// Developer experiment (not part of the library): variants of the 5x5 median TMA kernel, timed with CUDA events
// and compared bit for bit with the library kernel's structure (variant 0).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I video-watermarking-forgery-detection_b200/csrc \
//        tools/exp/m5exp.cu -o tools/exp/m5exp && tools/exp/m5exp [B H W]
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "median_pair_net.cuh"
#include "tma.cuh"
#include "wm_common.cuh"

namespace wm {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); }
int cuda_fail(cudaError_t e, const char* what) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return (int)e; }

__device__ __forceinline__ float feq(float a, float b) { float d; asm("set.eq.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ int first_match(float code, int n) { return n - 1 - ((__float_as_int(code) >> 23) - 127); }

// max(a, b) as a + b - min(a, b) on the bit patterns: two IMADs (FMA pipe) instead of one FMNMX (ALU pipe); the
// multipliers +1 / -1 come from kernel parameters so that ptxas cannot fold the IMAD back into an IADD3
struct CeIntSum {
    int one, neg1;
    __device__ __forceinline__ void operator()(float& a, float& b) const {
        const float lo = fminf(a, b);
        int s, h;
        asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(s) : "r"(__float_as_int(a)), "r"(one), "r"(__float_as_int(b)));
        asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(h) : "r"(__float_as_int(lo)), "r"(neg1), "r"(s));
        b = __int_as_float(h); a = lo;
    }
};
template <class CE> __device__ __forceinline__ void sort5c(float (&v)[5], const CE ce) {
    ce(v[0], v[1]); ce(v[3], v[4]); ce(v[2], v[4]); ce(v[2], v[3]); ce(v[0], v[3]);
    ce(v[0], v[2]); ce(v[1], v[4]); ce(v[1], v[3]); ce(v[1], v[2]);
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1,%2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 r;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rd));
    return r;
}

constexpr int M5_TW = 128, M5_TH = 72, M5_HALO = 4, M5_BW = M5_TW + 2 * M5_HALO, M5_BH = M5_TH + 4,
              M5_THREADS = 256, M5_ROWS = 36, M5_STAGES = 2, M5_STRIDE = ((M5_BW * M5_BH + 31) / 32) * 32;

struct MedTArgs { float* y; uint8_t* idx; int N, H, W, tiles_x, tiles_y; int64_t total; int one, neg1; };

template <bool I, class A, class B> struct pick { using type = B; };
template <class A, class B> struct pick<true, A, B> { using type = A; };

// VAR bits: 1 = sort5 with CeIntSum, 2 = merge10, 4 = mid6, 8 = FFMA2 pair search, 16 = predicated stores with 32-bit offsets
template <int VAR, bool WANT_IDX>
__global__ void __launch_bounds__(M5_THREADS, 2) m5_kernel(const __grid_constant__ CUtensorMap tmap, const MedTArgs a) {
    extern __shared__ __align__(128) float bufs[];
    __shared__ uint64_t full[M5_STAGES];
    const int tid = threadIdx.x;
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
#pragma unroll
        for (int s = 0; s < M5_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int per_plane = a.tiles_x * a.tiles_y;
    auto issue = [&](int64_t t, int s) {
        const int n = int(t / per_plane), rem = int(t - int64_t(n) * per_plane);
        const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
        mbar_expect_tx(&full[s], M5_BW * M5_BH * sizeof(float));
        tma_load_3d(bufs + s * M5_STRIDE, &tmap, tx * M5_TW - M5_HALO, ty * M5_TH - 2, n, &full[s]);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < M5_STAGES; ++s) {
            const int64_t t = int64_t(blockIdx.x) + int64_t(s) * gridDim.x;
            if (t < a.total) issue(t, s);
        }
    }
    const CeIntSum cei{a.one, a.neg1};
    const CeMinMax cem;
    const typename pick<(VAR & 1) != 0, CeIntSum, CeMinMax>::type ce_s = [&] { if constexpr ((VAR & 1) != 0) return cei; else return cem; }();
    const typename pick<(VAR & 2) != 0, CeIntSum, CeMinMax>::type ce_m = [&] { if constexpr ((VAR & 2) != 0) return cei; else return cem; }();
    const typename pick<(VAR & 4) != 0, CeIntSum, CeMinMax>::type ce_c = [&] { if constexpr ((VAR & 4) != 0) return cei; else return cem; }();
    const int c = tid & 127, strip = tid >> 7;
    int it = 0;
    for (int64_t t = blockIdx.x; t < a.total; t += gridDim.x, ++it) {
        const int s = it % M5_STAGES;
        mbar_wait(&full[s], (it / M5_STAGES) & 1);
        const int n = int(t / per_plane), rem = int(t - int64_t(n) * per_plane);
        const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
        const int gx = tx * M5_TW + c, gy0 = ty * M5_TH + strip * M5_ROWS;
        const float* col = bufs + s * M5_STRIDE + (strip * M5_ROWS) * M5_BW + M5_HALO - 2 + c;
        float srt[6][5], raw[WANT_IDX ? 6 : 1][5];
        auto load_row = [&](int row, int slot) {
            const float* p = col + row * M5_BW;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                srt[slot][k] = p[k];
                if (WANT_IDX) raw[slot][k] = srt[slot][k];
            }
            sort5c(srt[slot], ce_s);
        };
#pragma unroll
        for (int j = 0; j < 4; ++j) load_row(j, j);
        const bool col_ok = gx < a.W;
        const int64_t obase = (int64_t(n) * a.H + gy0) * a.W + gx;
        float* const yb = a.y + obase;
        uint8_t* const ib = WANT_IDX ? a.idx + obase : nullptr;
        const int rows_ok = col_ok ? a.H - gy0 : 0;       // output rows rr < rows_ok of this strip exist
        int off = 0;
        float mp[10];
#pragma unroll
        for (int k = 0; k < 5; ++k) { mp[k] = srt[1][k]; mp[5 + k] = srt[2][k]; }
        merge10_sorted_5_5(mp, ce_m);
#pragma unroll 1
        for (int r0 = 0; r0 < M5_ROWS; r0 += 6) {
            if ((VAR & 32) != 0 && gy0 + r0 >= a.H) break;      // rows below the image (warp-uniform)
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int r = r0 + 2 * u;
                load_row(r + 4, (2 * u + 4) % 6);
                load_row(r + 5, (2 * u + 5) % 6);
                float mq[10], v[20];
#pragma unroll
                for (int k = 0; k < 5; ++k) { mq[k] = srt[(2 * u + 3) % 6][k]; mq[5 + k] = srt[(2 * u + 4) % 6][k]; }
                merge10_sorted_5_5(mq, ce_m);
#pragma unroll
                for (int k = 0; k < 10; ++k) { v[k] = mp[k]; v[10 + k] = mq[k]; mp[k] = mq[k]; }
                mid6_of_2_sorted_10(v, ce_c);
                float med[2];
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    const int own = (2 * u + (o ? 5 : 0)) % 6;
                    float w[11];
#pragma unroll
                    for (int k = 0; k < 6; ++k) w[k] = v[7 + k];
#pragma unroll
                    for (int k = 0; k < 5; ++k) w[6 + k] = srt[own][k];
                    med[o] = median11_sorted_6_5(w);
                }
                int pos[2] = {0, 0};
                if (WANT_IDX) {
                    if constexpr ((VAR & 64) != 0) {
                        float2 rc[5];
                        const float2 two = make_float2(2.f, 2.f);
#pragma unroll
                        for (int j = 0; j < 5; ++j) {
                            rc[j] = make_float2(feq(raw[(2 * u + j) % 6][0], med[0]), feq(raw[(2 * u + 1 + j) % 6][0], med[1]));
#pragma unroll
                            for (int k = 1; k < 5; ++k)
                                rc[j] = ffma2(rc[j], two, make_float2(feq(raw[(2 * u + j) % 6][k], med[0]), feq(raw[(2 * u + 1 + j) % 6][k], med[1])));
                        }
                        // 25-bit code = ((((r0*32 + r1)*32 + r2)*32 + r3)*32 + r4); only a code of 25 set bits can round up to 2^25
                        float t0 = fmaf(fmaf(fmaf(fmaf(rc[0].x, 32.f, rc[1].x), 32.f, rc[2].x), 32.f, rc[3].x), 32.f, rc[4].x);
                        float t1 = fmaf(fmaf(fmaf(fmaf(rc[0].y, 32.f, rc[1].y), 32.f, rc[2].y), 32.f, rc[3].y), 32.f, rc[4].y);
                        pos[0] = first_match(fminf(t0, 33554430.f), 25);
                        pos[1] = first_match(fminf(t1, 33554430.f), 25);
                    } else if constexpr ((VAR & 128) != 0) {
                        // independent weighted sums (no Horner chain): 5 accumulators per output, immediate weights
#pragma unroll
                        for (int o = 0; o < 2; ++o) {
                            float rcs[5];
#pragma unroll
                            for (int j = 0; j < 5; ++j) {
                                rcs[j] = feq(raw[(2 * u + o + j) % 6][4], med[o]);
#pragma unroll
                                for (int k = 3; k >= 0; --k) rcs[j] = fmaf(feq(raw[(2 * u + o + j) % 6][k], med[o]), float(1 << (4 - k)), rcs[j]);
                            }
                            const float tt = fmaf(fmaf(fmaf(fmaf(rcs[0], 32.f, rcs[1]), 32.f, rcs[2]), 32.f, rcs[3]), 32.f, rcs[4]);
                            pos[o] = first_match(fminf(tt, 33554430.f), 25);
                        }
                    } else if constexpr ((VAR & 8) != 0) {
                        float2 hi = make_float2(0.f, 0.f), lo = make_float2(0.f, 0.f);
                        const float2 two = make_float2(2.f, 2.f);
#pragma unroll
                        for (int j = 0; j < 15; ++j)
                            hi = ffma2(hi, two, make_float2(feq(raw[(2 * u + j / 5) % 6][j % 5], med[0]),
                                                            feq(raw[(2 * u + 1 + j / 5) % 6][j % 5], med[1])));
#pragma unroll
                        for (int j = 15; j < 25; ++j)
                            lo = ffma2(lo, two, make_float2(feq(raw[(2 * u + j / 5) % 6][j % 5], med[0]),
                                                            feq(raw[(2 * u + 1 + j / 5) % 6][j % 5], med[1])));
                        pos[0] = hi.x != 0.f ? first_match(hi.x, 15) : 15 + first_match(lo.x, 10);
                        pos[1] = hi.y != 0.f ? first_match(hi.y, 15) : 15 + first_match(lo.y, 10);
                    } else {
#pragma unroll
                        for (int o = 0; o < 2; ++o) {
                            float hi = 0.f, lo = 0.f;
#pragma unroll
                            for (int j = 0; j < 15; ++j) hi = fmaf(hi, 2.f, feq(raw[(2 * u + o + j / 5) % 6][j % 5], med[o]));
#pragma unroll
                            for (int j = 15; j < 25; ++j) lo = fmaf(lo, 2.f, feq(raw[(2 * u + o + j / 5) % 6][j % 5], med[o]));
                            pos[o] = hi != 0.f ? first_match(hi, 15) : 15 + first_match(lo, 10);
                        }
                    }
                }
                if constexpr ((VAR & 16) != 0) {
#pragma unroll
                    for (int o = 0; o < 2; ++o) {
                        if (r + o < rows_ok) {
                            yb[off] = med[o];
                            if (WANT_IDX) ib[off] = (uint8_t)pos[o];
                        }
                        off += a.W;
                    }
                } else {
#pragma unroll
                    for (int o = 0; o < 2; ++o) {
                        const int rr = r + o;
                        if (col_ok && gy0 + rr < a.H) {
                            a.y[obase + int64_t(rr) * a.W] = med[o];
                            if (WANT_IDX) a.idx[obase + int64_t(rr) * a.W] = (uint8_t)pos[o];
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            const int64_t t2 = t + int64_t(M5_STAGES) * gridDim.x;
            if (t2 < a.total) issue(t2, s);
        }
    }
}
}  // namespace wm

using namespace wm;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int VAR, bool IDX>
static float run(const CUtensorMap& tm, MedTArgs ta, int reps) {
    const size_t smem = sizeof(float) * size_t(M5_STAGES) * M5_STRIDE;
    auto kern = m5_kernel<VAR, IDX>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t cap = int64_t(sm_count()) * 2;
    const unsigned grid = (unsigned)(ta.total < cap ? ta.total : cap);
    const int warm = getenv("M5_WARM") ? atoi(getenv("M5_WARM")) : 3;
    for (int i = 0; i < warm; ++i) kern<<<grid, M5_THREADS, smem>>>(tm, ta);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) kern<<<grid, M5_THREADS, smem>>>(tm, ta);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps * 1e3f;
}

int main(int argc, char** argv) {
    int B = 64, H = 512, W = 512;
    if (argc >= 4) { B = atoi(argv[1]); H = atoi(argv[2]); W = atoi(argv[3]); }
    const int N = B * 3;
    const size_t n = size_t(N) * H * W;
    std::vector<float> hx(n);
    uint32_t s = 12345u;
    for (size_t i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; hx[i] = ((s >> 8) & 0xffff) / 65535.f; if ((s >> 28) == 0) hx[i] = 0.f; }
    float *x, *y0, *y1; uint8_t *i0, *i1;
    CK(cudaMalloc(&x, n * 4)); CK(cudaMalloc(&y0, n * 4)); CK(cudaMalloc(&y1, n * 4)); CK(cudaMalloc(&i0, n)); CK(cudaMalloc(&i1, n));
    CK(cudaMemcpy(x, hx.data(), n * 4, cudaMemcpyHostToDevice));
    CUtensorMap tm;
    if (tmap_planes(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, N, H, W, int64_t(H) * W, W, M5_BW, M5_BH)) { fprintf(stderr, "tmap failed\n"); return 1; }
    MedTArgs ta{y0, i0, N, H, W, (W + M5_TW - 1) / M5_TW, (H + M5_TH - 1) / M5_TH, 0, 1, -1};
    ta.total = int64_t(N) * ta.tiles_x * ta.tiles_y;
    const int reps = getenv("M5_REPS") ? atoi(getenv("M5_REPS")) : 20;
    printf("shape %dx3x%dx%d\n", B, H, W);
    const float base = run<0, true>(tm, ta, reps);
    printf("VAR  0 idx : %8.1f us\n", base);
    std::vector<float> ry(n), ty_(n); std::vector<uint8_t> ri(n), ti(n);
    CK(cudaMemcpy(ry.data(), y0, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ri.data(), i0, n, cudaMemcpyDeviceToHost));
    ta.y = y1; ta.idx = i1;
    auto check = [&](int var, float us, bool idx) {
        CK(cudaMemcpy(ty_.data(), y1, n * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0, badi = 0;
        for (size_t i = 0; i < n; ++i) bad += memcmp(&ty_[i], &ry[i], 4) != 0;
        if (idx) { CK(cudaMemcpy(ti.data(), i1, n, cudaMemcpyDeviceToHost)); for (size_t i = 0; i < n; ++i) badi += ti[i] != ri[i]; }
        printf("VAR %2d %s: %8.1f us   mismatches y=%zu idx=%zu\n", var, idx ? "idx " : "noix", us, bad, badi);
        CK(cudaMemset(y1, 0xff, n * 4)); CK(cudaMemset(i1, 0xff, n));
    };
#define RUN(V) { float us = run<V, true>(tm, ta, reps); check(V, us, true); }
#define RUNN(V) { float us = run<V, false>(tm, ta, reps); check(V, us, false); }
    RUN(23) RUN(55) RUN(119) RUN(183) RUN(48) RUN(112) RUN(176) RUN(51) RUN(115) RUN(179) RUN(53) RUN(54) RUN(117)
    RUNN(0) RUNN(23) RUNN(55) RUNN(51) RUNN(48)
    return 0;
}
