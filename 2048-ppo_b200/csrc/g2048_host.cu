// g2048_host.cu -- error reporting, device checks.
#include "g2048_host.h"

namespace g2048 {

char* last_error_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

int num_sms() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

cudaError_t ensure_smem_impl(const void* kernel, int bytes) {
    struct Entry { const void* k; int dev; int bytes; };
    static thread_local Entry cache[64];
    static thread_local int used = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    for (int i = 0; i < used; ++i)
        if (cache[i].k == kernel && cache[i].dev == dev && cache[i].bytes >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && used < 64) cache[used++] = {kernel, dev, bytes};
    return e;
}

}  // namespace g2048

extern "C" {

const char* g2048_last_error(void) { return g2048::last_error_buf(); }
const char* g2048_version(void) { return "g2048 0.1 (sm_100a)"; }

int g2048_init(int device) {
    int count = 0;
    G2048_CHECK_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return g2048::fail(G2048_EINVAL, "device %d out of range (%d devices)", device, count);
    int major = 0;
    G2048_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) return g2048::fail(G2048_EARCH, "device %d has compute capability %d.x; this library is sm_100a only", device, major);
    return G2048_OK;      // a check only: the caller's current device is left alone (every entry point runs on the stream it is given)
}

}  // extern "C"
