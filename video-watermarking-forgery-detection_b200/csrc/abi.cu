// ABI bookkeeping: version, thread-local error message.
#include <stdarg.h>

#include "wm_common.cuh"

namespace wm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return (int)e;
}

static thread_local StoreEp g_ep = {nullptr, 0, 0};

void arm_store_epilogue(const StoreEp& e) { g_ep = e; }

StoreEp take_store_epilogue() {
    StoreEp e = g_ep;
    g_ep = StoreEp{nullptr, 0, 0, 0};
    return e;
}
bool reject_store_epilogue(const char* who) {
    if (!g_ep.x) return false;
    g_ep = StoreEp{nullptr, 0, 0, 0};
    set_error("%s: a store epilogue is armed but this entry point / code path does not apply one", who);
    return true;
}

}  // namespace wm

extern "C" int wm_set_store_epilogue(const float* x, int clamp01, int quantize) {
    WM_REQUIRE(x == nullptr || ::wm::aligned(x, 32), WM_E_ALIGN, "wm_set_store_epilogue: x must be 32-byte aligned");
    ::wm::arm_store_epilogue(::wm::StoreEp{x, clamp01, quantize, 0});
    return WM_OK;
}

extern "C" int wm_version(void) { return WM_ABI_VERSION; }
extern "C" const char* wm_last_error(void) { return wm::g_err; }
