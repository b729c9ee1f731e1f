// dh_api.cu -- version / error / device queries of libdynhor_b200.so
#include <string.h>

#include <vector>

#include "dh_common.h"

namespace dh {
char* err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}
}  // namespace dh

extern "C" {

int dh_version(void) { return 100; }  // 0.1.0

const char* dh_last_error(void) { return dh::err_buf(); }

int dh_struct_bytes(int32_t which) {
    switch (which) {
        case 0: return (int)sizeof(dh_sil);
        case 1: return (int)sizeof(dh_jointopt);
        case 2: return (int)sizeof(dh_corr);
        default: return dh::fail(DH_ERR_INVALID, "dh_struct_bytes: which must be 0, 1 or 2");
    }
}

int dh_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return dh::fail(DH_ERR_NO_DEVICE, "cudaGetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return dh::fail(DH_ERR_NO_DEVICE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return DH_OK;
}

int dh_dev_alloc(void** ptr, int64_t bytes) {
    DH_REQUIRE(ptr != nullptr && bytes > 0, "bad arguments");
    DH_CUDA(cudaMalloc(ptr, (size_t)bytes));
    DH_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
    return DH_OK;
}

int dh_dev_free(void* ptr) {
    if (ptr != nullptr) DH_CUDA(cudaFree(ptr));
    return DH_OK;
}

int dh_memcpy_d2d(void* dst, const void* src, int64_t bytes, void* stream) {
    DH_REQUIRE(dst && src && bytes > 0, "bad arguments");
    DH_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return DH_OK;
}

int dh_upload_rows(void* dst, const void* const* src_rows_host, int64_t row_bytes, int32_t n, void* stream) {
    DH_REQUIRE(dst && src_rows_host && row_bytes > 0 && n > 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    char* d = static_cast<char*>(dst);
    // one driver call for the whole batch where the runtime has it (CUDA >= 12.8; not on the legacy default stream)
    static thread_local int batch_ok = 1;
    if (batch_ok && st != nullptr && n > 1) {
        std::vector<void*> dsts(n), srcs(n);
        std::vector<size_t> sizes(n, (size_t)row_bytes);
        for (int i = 0; i < n; i++) {
            dsts[i] = d + (size_t)i * row_bytes;
            srcs[i] = const_cast<void*>(src_rows_host[i]);
        }
        cudaMemcpyAttributes attr;
        memset(&attr, 0, sizeof(attr));
        attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
        size_t attr_idx = 0, fail_idx = 0;
        cudaError_t e = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), (size_t)n, &attr, &attr_idx, 1,
                                             &fail_idx, st);
        if (e == cudaSuccess) return DH_OK;
        cudaGetLastError();   // not supported here: remember, and copy row by row
        batch_ok = 0;
    }
    for (int i = 0; i < n; i++)
        DH_CUDA(cudaMemcpyAsync(d + (size_t)i * row_bytes, src_rows_host[i], (size_t)row_bytes, cudaMemcpyHostToDevice, st));
    return DH_OK;
}

int dh_ipc_export(const void* ptr, void* handle64_host) {
    DH_REQUIRE(ptr && handle64_host, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DH_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64_host), const_cast<void*>(ptr)));
    return DH_OK;
}

int dh_ipc_open(const void* handle64_host, void** ptr) {
    DH_REQUIRE(ptr && handle64_host, "bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, sizeof(h));
    DH_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DH_OK;
}

int dh_ipc_close(void* ptr) {
    if (ptr != nullptr) DH_CUDA(cudaIpcCloseMemHandle(ptr));
    return DH_OK;
}

}  // extern "C"
