// dh_api.cu -- version / error / device queries of libdynhor_b200.so
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

}  // extern "C"
