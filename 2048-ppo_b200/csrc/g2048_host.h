// g2048_host.h -- host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../include/g2048.h"

namespace g2048 {

char* last_error_buf();   // thread-local, 512 bytes
int fail(int code, const char* fmt, ...);
int num_sms();            // SM count of the current device (cached per device)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device) instead of per launch.
// Keyed by the kernel's address (template instantiations share a function TYPE, not an address).
cudaError_t ensure_smem_impl(const void* kernel, int bytes);
template <class K>
inline cudaError_t ensure_smem(K kernel, int bytes) {
    return ensure_smem_impl(reinterpret_cast<const void*>(kernel), bytes);
}

#define G2048_CHECK_CUDA(expr)                                                                   \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return ::g2048::fail(G2048_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));  \
    } while (0)

#define G2048_CHECK_LAUNCH(name)                                                                 \
    do {                                                                                         \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e != cudaSuccess)                                                                   \
            return ::g2048::fail(G2048_ECUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
    } while (0)

#define G2048_REQUIRE(cond, msg)                                         \
    do {                                                                 \
        if (!(cond)) return ::g2048::fail(G2048_EINVAL, "%s", msg);     \
    } while (0)

}  // namespace g2048
