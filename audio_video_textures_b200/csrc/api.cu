// Error channel + small utilities of the C ABI (include/avtex.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void avtex_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int avtex_abi_version(void) { return AVTEX_ABI_VERSION; }

extern "C" const char *avtex_last_error(void) { return g_err; }

extern "C" int avtex_device_info(int device, int *sm_count, int *cc) {
    cudaDeviceProp prop;
    AVTEX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc) *cc = prop.major * 10 + prop.minor;
    return 0;
}

extern "C" int avtex_zero(void *ptr, int64_t bytes, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_CUDA(cudaMemsetAsync(ptr, 0, (size_t)bytes, as_stream(stream)));
    return 0;
}
