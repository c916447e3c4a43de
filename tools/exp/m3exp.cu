// Developer experiment: variants of the 3x3 median TMA kernel (arg-median plane on).  See m5exp.cu.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "tma.cuh"
#include "wm_common.cuh"

namespace wm {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); }
int cuda_fail(cudaError_t e, const char* what) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return (int)e; }
__device__ __forceinline__ float feq(float a, float b) { float d; asm("set.eq.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ int first_match(float code, int n) { return n - 1 - ((__float_as_int(code) >> 23) - 127); }

struct Pm { int one, neg1; };
template <bool INT>
__device__ __forceinline__ float mid3(float a, float b, float c, float lo, float hi, const Pm pm) {
    if constexpr (INT) {
        int s;
        asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(s) : "r"(__float_as_int(a)), "r"(pm.one), "r"(__float_as_int(b)));
        asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(s) : "r"(__float_as_int(c)), "r"(pm.one));
        asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(s) : "r"(__float_as_int(lo)), "r"(pm.neg1));
        asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(s) : "r"(__float_as_int(hi)), "r"(pm.neg1));
        return __int_as_float(s);
    } else {
        return __int_as_float(__float_as_int(a) ^ __float_as_int(b) ^ __float_as_int(c) ^ __float_as_int(lo) ^ __float_as_int(hi));
    }
}

constexpr int MT_TW = 128, MT_TH = 64, MT_HALO = 4, MT_BW = MT_TW + 2 * MT_HALO, MT_BH = MT_TH + 2,
              MT_THREADS = 256, MT_ROWS = 8, MT_STRIDE = ((MT_BW * MT_BH + 31) / 32) * 32;
struct MedTArgs { float* y; uint8_t* idx; int N, H, W, tiles_x, tiles_y; int64_t total; Pm pm; };

// VAR bits: 1 = row-triple mid on FMA pipe, 2 = column mid-of-mids, 4 = final med3, 8 = 3 independent search chains,
//           16 = 2-stage ring and 3 CTAs per SM (<= 85 registers) instead of 3 stages and 2 CTAs
template <int VAR, bool WANT_IDX>
__global__ void __launch_bounds__(MT_THREADS, (VAR & 16) ? 3 : 2) m3_kernel(const __grid_constant__ CUtensorMap tmap, const MedTArgs a) {
    constexpr int MT_STAGES = (VAR & 16) ? 2 : 3;
    extern __shared__ __align__(128) float bufs[];
    __shared__ uint64_t full[MT_STAGES];
    const int tid = threadIdx.x;
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
#pragma unroll
        for (int s = 0; s < MT_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int per_plane = a.tiles_x * a.tiles_y;
    auto issue = [&](int64_t t, int s) {
        const int n = int(t / per_plane), rem = int(t - int64_t(n) * per_plane);
        const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
        mbar_expect_tx(&full[s], MT_BW * MT_BH * sizeof(float));
        tma_load_3d(bufs + s * MT_STRIDE, &tmap, tx * MT_TW - MT_HALO, ty * MT_TH - 1, n, &full[s]);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < MT_STAGES; ++s) {
            const int64_t t = int64_t(blockIdx.x) + int64_t(s) * gridDim.x;
            if (t < a.total) issue(t, s);
        }
    }
    const Pm pm = a.pm;
    const int cg = tid & 31, strip = tid >> 5;
    int it = 0;
    for (int64_t t = blockIdx.x; t < a.total; t += gridDim.x, ++it) {
        const int s = it % MT_STAGES;
        mbar_wait(&full[s], (it / MT_STAGES) & 1);
        const int n = int(t / per_plane), rem = int(t - int64_t(n) * per_plane);
        const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
        const int gx = tx * MT_TW + 4 * cg, gy0 = ty * MT_TH + strip * MT_ROWS;
        const float* col = bufs + s * MT_STRIDE + (strip * MT_ROWS) * MT_BW + MT_HALO + 4 * cg;
        float raw[3][6], lo[3][4], mi[3][4], hi[3][4];
        auto load_row = [&](int row, int slot) {
            const float* p = col + row * MT_BW;
            const float4 c = *reinterpret_cast<const float4*>(p);
            raw[slot][0] = p[-1]; raw[slot][1] = c.x; raw[slot][2] = c.y; raw[slot][3] = c.z; raw[slot][4] = c.w;
            raw[slot][5] = p[4];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float u = raw[slot][c4], v = raw[slot][c4 + 1], w = raw[slot][c4 + 2];
                const float l = fmin3(u, v, w), h = fmax3(u, v, w);
                lo[slot][c4] = l; hi[slot][c4] = h; mi[slot][c4] = mid3<(VAR & 1) != 0>(u, v, w, l, h, pm);
            }
        };
        load_row(0, 0);
        load_row(1, 1);
        const bool col_ok = gx < a.W;
        const int64_t obase = (int64_t(n) * a.H + gy0) * a.W + gx;
#pragma unroll
        for (int r = 0; r < MT_ROWS; ++r) {
            load_row(r + 2, (r + 2) % 3);
            float4 o;
            float* op = &o.x;
            uint32_t packed = 0;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float m0 = mi[0][c4], m1 = mi[1][c4], m2 = mi[2][c4];
                const float A = fmax3(lo[0][c4], lo[1][c4], lo[2][c4]);
                const float B = mid3<(VAR & 2) != 0>(m0, m1, m2, fmin3(m0, m1, m2), fmax3(m0, m1, m2), pm);
                const float C = fmin3(hi[0][c4], hi[1][c4], hi[2][c4]);
                const float med = mid3<(VAR & 4) != 0>(A, B, C, fmin3(A, B, C), fmax3(A, B, C), pm);
                op[c4] = med;
                if (WANT_IDX) {
                    if constexpr ((VAR & 8) != 0) {
                        float cr[3];
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            const float* rw = raw[(r + j) % 3] + c4;
                            cr[j] = fmaf(feq(rw[0], med), 4.f, fmaf(feq(rw[1], med), 2.f, feq(rw[2], med)));
                        }
                        const float code = fmaf(fmaf(cr[0], 8.f, cr[1]), 8.f, cr[2]);
                        packed |= uint32_t(first_match(code, 9)) << (8 * c4);
                    } else {
                        float code = 0.f;
#pragma unroll
                        for (int j = 0; j < 9; ++j) code = fmaf(code, 2.f, feq(raw[(r + j / 3) % 3][c4 + j % 3], med));
                        packed |= uint32_t(first_match(code, 9)) << (8 * c4);
                    }
                }
            }
            if (col_ok && gy0 + r < a.H) {
                stg128(a.y + obase + int64_t(r) * a.W, o);
                if (WANT_IDX) *reinterpret_cast<uint32_t*>(a.idx + obase + int64_t(r) * a.W) = packed;
            }
        }
        __syncthreads();
        if (tid == 0) {
            const int64_t t2 = t + int64_t(MT_STAGES) * gridDim.x;
            if (t2 < a.total) issue(t2, s);
        }
    }
}
}  // namespace wm

using namespace wm;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int VAR, bool IDX>
static float run(const CUtensorMap& tm, MedTArgs ta, int reps) {
    constexpr int MT_STAGES = (VAR & 16) ? 2 : 3;
    const size_t smem = sizeof(float) * size_t(MT_STAGES) * MT_STRIDE;
    auto kern = m3_kernel<VAR, IDX>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t cap = int64_t(sm_count()) * ((VAR & 16) ? 3 : 2);
    const unsigned grid = (unsigned)(ta.total < cap ? ta.total : cap);
    const int warm = getenv("M5_WARM") ? atoi(getenv("M5_WARM")) : 3;
    for (int i = 0; i < warm; ++i) kern<<<grid, MT_THREADS, smem>>>(tm, ta);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) kern<<<grid, MT_THREADS, smem>>>(tm, ta);
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
    if (tmap_planes(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, N, H, W, int64_t(H) * W, W, MT_BW, MT_BH)) { fprintf(stderr, "tmap failed\n"); return 1; }
    MedTArgs ta{y0, i0, N, H, W, (W + MT_TW - 1) / MT_TW, (H + MT_TH - 1) / MT_TH, 0, Pm{1, -1}};
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
    RUN(6) RUN(16) RUN(22) RUN(23)
    RUNN(0) RUNN(16)
    return 0;
}
