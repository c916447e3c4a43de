// Probe: which starting coordinates / box widths does a tiled TMA load of 2-byte elements accept?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdarg>
#include "tma.cuh"
namespace wm {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); }
int cuda_fail(cudaError_t e, const char* what) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return (int)e; }
__global__ void probe(const __grid_constant__ CUtensorMap tm, int c0, int c1, int bytes, unsigned* out) {
    extern __shared__ __align__(128) unsigned char buf[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(&bar, bytes); tma_load_3d(buf, &tm, c0, c1, 0, &bar); }
    mbar_wait(&bar, 0);
    unsigned s = 0;
    for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) s += reinterpret_cast<const uint16_t*>(buf)[i];
    atomicAdd(out, s);
}
}
using namespace wm;
int main() {
    const int W = 136, H = 80, N = 2;
    std::vector<uint16_t> h(size_t(N) * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = 1;
    uint16_t* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    unsigned* out; cudaMalloc(&out, 4);
    const int boxes[][2] = {{136, 66}, {144, 66}, {128, 66}, {136, 8}};
    const int starts[] = {0, 8, -8, 4, -4, 2, 1};
    for (auto& b : boxes) {
        CUtensorMap tm;
        int rc = tmap_planes(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, N, H, W, int64_t(H) * W, W, b[0], b[1]);
        printf("box %dx%d encode rc=%d\n", b[0], b[1], rc);
        if (rc) continue;
        for (int c0 : starts) {
            cudaMemset(out, 0, 4);
            const int bytes = b[0] * b[1] * 2;
            cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            probe<<<1, 128, bytes>>>(tm, c0, -1, bytes, out);
            cudaError_t e = cudaDeviceSynchronize();
            unsigned v = 0; if (e == cudaSuccess) cudaMemcpy(&v, out, 4, cudaMemcpyDeviceToHost);
            printf("   start x=%3d y=-1: %s  sum=%u\n", c0, cudaGetErrorString(e), v);
            if (e != cudaSuccess) { printf("   (context lost)\n"); return 0; }
        }
    }
    return 0;
}
