// Developer experiment (not part of the library): 5x5 median BACKWARD as a conflict-free scatter in shared memory.
//
// The library kernel is a gather: every input position tests the 25 outputs whose window holds it (25 FSET + 12.5
// FFMA2 + register moves per value, ALU-pipe bound at 0.56 of the HBM roofline).  Each output contributes to exactly
// ONE position, so a scatter does 1/25 of the tests - but two outputs may hit the same position.  Outputs whose
// coordinates agree modulo 5 in both axes have disjoint 5x5 windows, so the 25 "colours" (row mod 5, col mod 5) can
// be scattered one colour at a time with plain shared-memory read-add-write, a barrier between colours, no atomics,
// and a fixed summation order (colour order): deterministic.
//   thread (warp w, lane l < 27) owns the 5x5 block of outputs at rows 5w.., cols 5l.. of the tile's 40 x 135 output
//   region; phase (a, b) adds its element (a, b) into the accumulator tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I video-watermarking-forgery-detection_b200/csrc \
//        tools/exp/m5bexp.cu -o tools/exp/m5bexp -ldl && tools/exp/m5bexp [B H W]
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>
#include "tma.cuh"
#include "wm_common.cuh"

namespace wm {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); }
int cuda_fail(cudaError_t e, const char* what) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return (int)e; }

constexpr int S5_TW = 128, S5_TH = 36, S5_THREADS = 256;
constexpr int S5_QH = S5_TH + 4;                    // output rows of a tile: 8 warps x 5
constexpr int S5_GW = 140;                          // cotangent box: cols x0-4 .. x0+136
constexpr int S5_IW = 160;                          // idx box: cols x0-16 .. x0+144
constexpr int S5_AH = S5_TH + 8;                    // accumulator: cols x0-4 .. (S5_AW floats per row), rows y0-4 ..
constexpr int S5_BX = 27;                           // 5-column blocks per tile row (lanes in use)

struct ScArgs { float* gx; int N, H, W, tiles_x, tiles_y; int64_t total; };

// VAR bit 1: idx offsets through a 25-entry shared table instead of arithmetic; 2: idx rows as two aligned words + PRMT;
// 32: block-wide barrier only when the ROW colour changes (for one row colour the warps' target rows are disjoint: the
// five column colours need warp-level ordering only);  timing only (wrong results): 4 = no scatter phases,
// 8 = no store / clear pass, 16 = no preload.   S5_AW: accumulator row stride (160: rows on the same banks)
template <int VAR, int MINB, int S5_AW>
__global__ void __launch_bounds__(S5_THREADS, MINB) m5b_scatter_kernel(const __grid_constant__ CUtensorMap tm_g,
                                                                      const __grid_constant__ CUtensorMap tm_i, const ScArgs a) {
    extern __shared__ __align__(128) float sm[];
    float* gbox = sm;                                                   // [QH][GW]
    float* acc = gbox + S5_QH * S5_GW;                                  // [AH][AW]
    uint8_t* ibox = reinterpret_cast<uint8_t*>(acc + S5_AH * S5_AW);    // [QH][IW]
    __shared__ uint64_t full;
    __shared__ int lut[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 32) { const int iy = tid / 5; lut[tid] = tid < 25 ? iy * S5_AW + (tid - 5 * iy) : 0; }
    if (tid == 0) {
        tma_prefetch_desc(&tm_g); tma_prefetch_desc(&tm_i);
        mbar_init(&full, 1);
        mbar_fence_init();
    }
    for (int i = tid; i < S5_AH * S5_AW / 4; i += S5_THREADS) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int per_plane = a.tiles_x * a.tiles_y;
    auto issue = [&](int64_t t) {
        const int n = int(t / per_plane), rem = int(t - int64_t(n) * per_plane);
        const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
        mbar_expect_tx(&full, S5_GW * S5_QH * sizeof(float) + S5_IW * S5_QH);
        tma_load_3d(gbox, &tm_g, tx * S5_TW - 4, ty * S5_TH - 2, n, &full);
        tma_load_3d(ibox, &tm_i, tx * S5_TW - 16, ty * S5_TH - 2, n, &full);
    };
    if (tid == 0 && blockIdx.x < a.total) issue(blockIdx.x);
    const bool act = lane < S5_BX;
    const int lq = act ? lane : 0;
    const float* gsrc = gbox + (5 * warp) * S5_GW + 5 * lq + 2;          // output col c <-> box col c + 2
    const uint8_t* isrc = ibox + (5 * warp) * S5_IW + 5 * lq + 14;       //              <-> idx box col c + 14
    float* abase = acc + (5 * warp) * S5_AW + 5 * lq;                    // window position (iy, ix) of output (r, c): acc[r + iy][c + ix]
    int it = 0;
    for (int64_t t = blockIdx.x; t < a.total; t += gridDim.x, ++it) {
        mbar_wait(&full, it & 1);
        float g[5][5]; int off[5][5];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            uint32_t lo5 = 0, hi5 = 0;         // the row's 5 idx bytes: lo5 = bytes 0..3, hi5 = byte 4
            if (VAR & 2) {
                const uint32_t boff = uint32_t(isrc - ibox) + r * S5_IW;
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(ibox + (boff & ~3u));
                const uint32_t w0 = wp[0], w1 = wp[1], k = boff & 3u;
                lo5 = __byte_perm(w0, w1, 0x3210u + 0x1111u * k);
                hi5 = __byte_perm(w0, w1, 0x4444u + 0x1111u * k) & 0xffu;
            }
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                g[r][c] = (VAR & 16) ? float(r + c + lane) : gsrc[r * S5_GW + c];
                const int raw = (VAR & 16) ? ((r * 5 + c + lane) % 25) : (VAR & 2) ? int(c < 4 ? (lo5 >> (8 * c)) & 0xffu : hi5) : int(isrc[r * S5_IW + c]);
                const int id = min(raw, 24);
                if (VAR & 1) off[r][c] = lut[id] + r * S5_AW + c;
                else { const int iy = (id * 52) >> 8; off[r][c] = iy * (S5_AW - 5) + id + r * S5_AW + c; }
            }
        }
        __syncthreads();                    // every thread holds its 25 outputs: the boxes may be overwritten
        if (tid == 0 && t + gridDim.x < a.total) issue(t + gridDim.x);
        if (!(VAR & 4)) {
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    if (act) abase[off[r][c]] += g[r][c];
                    if ((VAR & 32) && c < 4) __syncwarp(); else __syncthreads();
                }
        } else {
            float sacc = 0.f; int so = 0;
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int c = 0; c < 5; ++c) { sacc += g[r][c]; so += off[r][c]; }
            if (sacc == 123.456f && so == 77) abase[0] = 1.f;
            __syncthreads();
        }
        // store the tile's 36 x 128 positions and clear the accumulator
        const int n = int(t / per_plane), rem = int(t - int64_t(n) * per_plane);
        const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
        const int gx = tx * S5_TW + 4 * lane;
        float* dst = a.gx + (int64_t(n) * a.H + ty * S5_TH) * a.W + gx;
#pragma unroll
        for (int k = 0; k < ((VAR & 8) ? 1 : 5); ++k) {
            const int row = warp + 8 * k;
            if (row < S5_TH) {
                const float4 v = *reinterpret_cast<const float4*>(acc + (row + 4) * S5_AW + 4 + 4 * lane);
                if (gx < a.W && ty * S5_TH + row < a.H) stg128(dst + int64_t(row) * a.W, v);
            }
        }
        __syncthreads();
        if (!(VAR & 8))
            for (int i = tid; i < S5_AH * S5_AW / 4; i += S5_THREADS) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
    }
}

// reference: plain gather in the library kernel's summation order
template <int K>
__global__ void m5b_ref_kernel(const float* __restrict__ gy, const uint8_t* __restrict__ idx, int64_t idx_sh, float* __restrict__ gx,
                               int N, int H, int W) {
    constexpr int R = K / 2;
    const int64_t total = int64_t(N) * H * W;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int w = int(i % W), h = int((i / W) % H); const int64_t n = i / (int64_t(H) * W);
        float acc = 0.f;
        for (int dy = -R; dy <= R; ++dy)
            for (int dx = -R; dx <= R; ++dx) {
                const int qy = h + dy, qx = w + dx;
                if (qy < 0 || qy >= H || qx < 0 || qx >= W) continue;
                if (idx[(n * H + qy) * idx_sh + qx] == (R - dy) * K + (R - dx)) acc += gy[(n * H + qy) * W + qx];
            }
        gx[i] = acc;
    }
}
}  // namespace wm
using namespace wm;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int VAR, int MINB, int S5_AW = 144>
static float run(const CUtensorMap& tg, const CUtensorMap& ti, ScArgs a, int reps) {
    const size_t smem = sizeof(float) * (S5_QH * S5_GW + S5_AH * S5_AW) + S5_QH * S5_IW;
    auto kern = m5b_scatter_kernel<VAR, MINB, S5_AW>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, S5_THREADS, smem));
    const int64_t cap = int64_t(sm_count()) * blocks;
    const unsigned grid = (unsigned)(a.total < cap ? a.total : cap);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) kern<<<grid, S5_THREADS, smem>>>(tg, ti, a);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) kern<<<grid, S5_THREADS, smem>>>(tg, ti, a);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("  [%d CTAs/SM, smem %zu B] ", blocks, smem);
    return ms * 1000.f / reps;
}

int main(int argc, char** argv) {
    int B = 64, H = 512, W = 512;
    if (argc >= 4) { B = atoi(argv[1]); H = atoi(argv[2]); W = atoi(argv[3]); }
    const int N = B * 3;
    const size_t n = size_t(N) * H * W;
    std::vector<float> hg(n); std::vector<uint8_t> hi(n);
    uint32_t s = 12345u;
    for (size_t i = 0; i < n; ++i) {
        s = s * 1664525u + 1013904223u; hg[i] = ((s >> 8) & 0xffff) / 65535.f - 0.5f;
        s = s * 1664525u + 1013904223u; hi[i] = (s >> 10) % 25;
        if (((s >> 20) & 7) == 0) hi[i] = 12;            // natural images: the centre is the median more often
    }
    float *gy, *g0, *g1; uint8_t* idx;
    CK(cudaMalloc(&gy, n * 4)); CK(cudaMalloc(&g0, n * 4)); CK(cudaMalloc(&g1, n * 4)); CK(cudaMalloc(&idx, n));
    CK(cudaMemcpy(gy, hg.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(idx, hi.data(), n, cudaMemcpyHostToDevice));
    printf("shape %dx3x%dx%d\n", B, H, W);
    m5b_ref_kernel<5><<<148 * 8, 256>>>(gy, idx, W, g0, N, H, W);
    CK(cudaDeviceSynchronize());
    std::vector<float> r0(n), r1(n);
    CK(cudaMemcpy(r0.data(), g0, n * 4, cudaMemcpyDeviceToHost));

    // library kernels (WM_LIBS = colon-separated paths; default: the built library)
    std::string libs = getenv("WM_LIBS") ? getenv("WM_LIBS") : "video-watermarking-forgery-detection_b200/wmattack/libwmattack.so";
    std::vector<uint8_t> hi3(n); for (size_t i = 0; i < n; ++i) hi3[i] = hi[i] % 9;
    uint8_t* idx3; CK(cudaMalloc(&idx3, n)); CK(cudaMemcpy(idx3, hi3.data(), n, cudaMemcpyHostToDevice));
    float* g3; CK(cudaMalloc(&g3, n * 4));
    m5b_ref_kernel<3><<<148 * 8, 256>>>(gy, idx3, W, g3, N, H, W);
    CK(cudaDeviceSynchronize());
    std::vector<float> r3(n); CK(cudaMemcpy(r3.data(), g3, n * 4, cudaMemcpyDeviceToHost));
    for (size_t pos = 0; pos < libs.size();) {
        size_t e = libs.find(':', pos); if (e == std::string::npos) e = libs.size();
        const std::string path = libs.substr(pos, e - pos); pos = e + 1;
        void* lib = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!lib) { printf("%s: %s\n", path.c_str(), dlerror()); continue; }
        using bwd_t = int (*)(const float*, const uint8_t*, int64_t, float*, int, int, int, int, void*);
        auto f = (bwd_t)dlsym(lib, "wm_median_bwd");
        for (int k = 3; k <= 5; k += 2) {
            const uint8_t* ip = k == 3 ? idx3 : idx; const std::vector<float>& ref = k == 3 ? r3 : r0;
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            CK(cudaMemset(g1, 0xff, n * 4));
            for (int i = 0; i < 3; ++i) f(gy, ip, W, g1, N, H, W, k, nullptr);
            CK(cudaEventRecord(e0));
            for (int i = 0; i < 20; ++i) f(gy, ip, W, g1, N, H, W, k, nullptr);
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            CK(cudaMemcpy(r1.data(), g1, n * 4, cudaMemcpyDeviceToHost));
            size_t bad = 0; for (size_t i = 0; i < n; ++i) bad += memcmp(&ref[i], &r1[i], 4) != 0;
            printf("%-40s k=%d : %8.1f us   bit mismatches vs plain gather %zu\n", path.substr(path.rfind('/') + 1).c_str(), k, ms * 1000.f / 20, bad);
        }
    }
    if (getenv("WM_LIBS")) return 0;

    CUtensorMap tg, ti;
    if (tmap_planes(&tg, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, gy, N, H, W, int64_t(H) * W, W, S5_GW, S5_QH)) { fprintf(stderr, "tmap g failed\n"); return 1; }
    if (tmap_planes(&ti, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, idx, N, H, W, int64_t(H) * W, W, S5_IW, S5_QH)) { fprintf(stderr, "tmap i failed\n"); return 1; }
    ScArgs a{g1, N, H, W, (W + S5_TW - 1) / S5_TW, (H + S5_TH - 1) / S5_TH, 0};
    a.total = int64_t(N) * a.tiles_x * a.tiles_y;
    auto check = [&](const char* name, float us) {
        CK(cudaMemcpy(r1.data(), g1, n * 4, cudaMemcpyDeviceToHost));
        size_t bits = 0; double worst = 0;
        for (size_t i = 0; i < n; ++i) { bits += memcmp(&r0[i], &r1[i], 4) != 0; const double d = fabs(double(r0[i]) - r1[i]); if (!(d <= worst)) worst = d; }
        printf("%s: %8.1f us   max |diff| vs gather %.3g (%zu values differ in the last bits: summation order)\n", name, us, worst, bits);
        CK(cudaMemset(g1, 0xff, n * 4));
    };
    CK(cudaMemset(g1, 0xff, n * 4));
    { float us = run<0, 3>(tg, ti, a, 20); check("scatter arith, 3/SM", us); }
    { float us = run<1, 3>(tg, ti, a, 20); check("scatter lut,   3/SM", us); }
    { float us = run<2, 3>(tg, ti, a, 20); check("scatter arith, word idx", us); }
    { float us = run<3, 3>(tg, ti, a, 20); check("scatter lut,   word idx", us); }
    { float us = run<3, 3, 160>(tg, ti, a, 20); check("lut, word idx, AW 160  ", us); }
    { float us = run<3 + 32, 3>(tg, ti, a, 20); check("lut, word idx, syncwarp", us); }
    { float us = run<3 + 32, 3, 160>(tg, ti, a, 20); check("lut, word, syncwarp 160", us); }
    { float us = run<2 + 32, 3, 160>(tg, ti, a, 20); check("arith,word,syncwarp 160", us); }
    { float us = run<3 + 32, 2, 160>(tg, ti, a, 20); check("lut, word, syncwarp 160 2/SM", us); }
    { float us = run<2 + 4, 3>(tg, ti, a, 20); check("TIMING no phases       ", us); }
    { float us = run<2 + 8, 3>(tg, ti, a, 20); check("TIMING no store/clear  ", us); }
    { float us = run<2 + 4 + 8, 3>(tg, ti, a, 20); check("TIMING preload only    ", us); }
    { float us = run<16 + 4 + 8, 3>(tg, ti, a, 20); check("TIMING TMA only        ", us); }
    { float us = run<16, 3>(tg, ti, a, 20); check("TIMING no preload      ", us); }
    // determinism: two runs, bit-identical
    run<0, 3>(tg, ti, a, 1); CK(cudaMemcpy(r0.data(), g1, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemset(g1, 0xff, n * 4));
    run<0, 3>(tg, ti, a, 1); CK(cudaMemcpy(r1.data(), g1, n * 4, cudaMemcpyDeviceToHost));
    printf("\nrun-to-run: %s\n", memcmp(r0.data(), r1.data(), n * 4) == 0 ? "bit-identical" : "DIFFERENT");
    return 0;
}
