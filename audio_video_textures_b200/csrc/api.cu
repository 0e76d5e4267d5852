// Error channel + small utilities of the C ABI (include/avtex.h).
#include <stdarg.h>
#include <string.h>

#include <atomic>

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

// cudaGetDeviceProperties costs milliseconds; the three attributes needed are cached per device.
extern "C" int avtex_device_info(int device, int *sm_count, int *cc) {
    // packed (sms << 16 | cc) + 1 in one atomic word per device: a racing second query stores the same value
    static std::atomic<int> cached[64];
    AVTEX_REQUIRE(device >= 0 && device < 64, "device_info: device index %d out of range", device);
    int packed = cached[device].load(std::memory_order_acquire);
    if (packed == 0) {
        int sms = 0, major = 0, minor = 0;
        AVTEX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        AVTEX_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
        AVTEX_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
        packed = ((sms << 16) | (major * 10 + minor)) + 1;
        cached[device].store(packed, std::memory_order_release);
    }
    if (sm_count) *sm_count = (packed - 1) >> 16;
    if (cc) *cc = (packed - 1) & 0xffff;
    return 0;
}

extern "C" int avtex_zero(void *ptr, int64_t bytes, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_CUDA(cudaMemsetAsync(ptr, 0, (size_t)bytes, as_stream(stream)));
    return 0;
}
