// dh_common.h -- error plumbing shared by the translation units of libdynhor_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/dynhor_b200.h"

namespace dh {

char* err_buf();  // thread-local, 512 bytes (dh_api.cu)

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

#define DH_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return ::dh::fail(DH_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,                 \
                              cudaGetErrorString(e__));                                                 \
    } while (0)

#define DH_LAUNCH_OK(name)                                                                              \
    do {                                                                                                \
        cudaError_t e__ = cudaGetLastError();                                                           \
        if (e__ != cudaSuccess)                                                                         \
            return ::dh::fail(DH_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__));   \
    } while (0)

#define DH_REQUIRE(cond, ...)                                                                           \
    do {                                                                                                \
        if (!(cond)) return ::dh::fail(DH_ERR_INVALID, __VA_ARGS__);                                    \
    } while (0)

// dh_corr.cu: memset of the partial sums + the streaming correspondence kernel, stream-ordered
int launch_corr(const float* records, int B, int C, const float* Rmat, const float* trans, const float* scale,
                const float* K, int S, float delta, float* partials, int nslots, cudaStream_t st);

}  // namespace dh
