// dh_dino.cu -- DINO template matching (pose_initializtion.py:295-311): bf16 GEMM + top-k.  (under construction)
#include "dh_common.h"

extern "C" {

int dh_dino_workspace_bytes(int32_t N, int32_t Fm, int64_t Kdim, int64_t* bytes) {
    DH_REQUIRE(bytes != nullptr && N > 0 && Fm > 0 && Kdim > 0, "bad arguments");
    *bytes = 0;
    return DH_OK;
}

int dh_dino_topk(const void*, const void*, int32_t, int32_t, int64_t, int32_t, float*, float*, int32_t*, void*,
                 int64_t, void*) {
    return dh::fail(DH_ERR_UNSUPPORTED, "dh_dino_topk: tcgen05 kernel not built yet");
}

int dh_dino_prescale(const float*, const float*, int32_t, int32_t, int32_t, void*, void*) {
    return dh::fail(DH_ERR_UNSUPPORTED, "dh_dino_prescale: not built yet");
}

}  // extern "C"
