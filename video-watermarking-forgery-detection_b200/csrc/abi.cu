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

}  // namespace wm

extern "C" int wm_version(void) { return WM_ABI_VERSION; }
extern "C" const char* wm_last_error(void) { return wm::g_err; }
