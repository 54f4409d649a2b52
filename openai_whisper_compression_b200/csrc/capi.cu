// capi.cu -- error plumbing and device queries of the C ABI.
#include "common.cuh"

#include <mutex>

namespace {
thread_local char g_err[512] = "";
}

void wq_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int wq_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return WQ_SM_COUNT_FALLBACK;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = WQ_SM_COUNT_FALLBACK;
        cached[dev] = n;
    }
    return cached[dev];
}

int wq_check_device() {
    int dev = 0, major = 0;
    WQ_CUDA(cudaGetDevice(&dev));
    WQ_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
        wq_set_error("libwhisperq is built for sm_100a only; device %d has compute capability %d.x", dev, major);
        return WQ_ERR_UNSUPPORTED;
    }
    return WQ_OK;
}

extern "C" const char *wq_last_error(void) { return g_err; }

extern "C" int wq_version(void) { return 100; /* 0.1.0 */ }

extern "C" int wq_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    WQ_CUDA(cudaGetDevice(&dev));
    if (sm_count) WQ_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (cc_major) WQ_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_minor) WQ_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    return WQ_OK;
}
