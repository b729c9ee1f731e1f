// dh_roi.cu -- ROI preprocessing that builds the joint optimisation's target masks, all frames at once.
//
// Replaces the per-frame CPU loop of ObjTracker/run.py:26-72 (`process_input`): tight bounding box of the object
// mask (+5 px), square box x1.3 (utils/bbox.py:73-89), ROIAlign crops to S x S of the object mask, the hand
// (occluder) mask and the RGB image (utils/bbox.py:8-36, detectron2 BitMasks.crop_and_resize), and the tri-state
// target mask 1 / 0 / -1 (utils/maskutils.py:8-30) that jointopt.py:50-53 turns into ref / keep masks.
// Arithmetic: dh_roi_core.h (bit-exact against torchvision's CPU roi_align, which is what detectron2's ROIAlign
// calls).  Compiled with --fmad=false.
//   k_roi_init    bounds <- (+inf, -1, +inf, -1)
//   k_roi_bounds  one pass over the object bit masks (16-byte loads): min / max row and column per frame
//   k_roi_crop    one thread per output cell: boxes, adaptive-grid bilinear samples of mask / occluder / image
#include "dh_common.h"
#include "dh_roi_core.h"

namespace {

using namespace dh;

constexpr int kThreads = 256;
constexpr int kRowsPerCta = 8;

__global__ void k_roi_init(int32_t* __restrict__ bounds, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 4 * B) bounds[i] = (i & 1) ? -1 : 0x7fffffff;   // min_row, max_row, min_col, max_col
}

__global__ void __launch_bounds__(kThreads)
k_roi_bounds(const uint8_t* __restrict__ obj_bits, int H, int W, int32_t* __restrict__ bounds) {
    const int b = blockIdx.y, row0 = blockIdx.x * kRowsPerCta;
    const int rows = min(kRowsPerCta, H - row0);
    const uint8_t* base = obj_bits + ((size_t)b * H + row0) * W;
    int rmin = 0x7fffffff, rmax = -1, cmin = 0x7fffffff, cmax = -1;
    const size_t nbytes = (size_t)rows * W;
    if ((W & 15) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0) {
        const uint4* v = reinterpret_cast<const uint4*>(base);
        const int wpr = W >> 4;   // 16-byte chunks per row
        for (int i = threadIdx.x; i < rows * wpr; i += kThreads) {
            const uint4 q = __ldg(v + i);
            if ((q.x | q.y | q.z | q.w) == 0u) continue;
            const int r = i / wpr, c0 = (i - r * wpr) << 4;
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (w[k]) {
                    cmin = min(cmin, c0 + 4 * k + ((__ffs((int)w[k]) - 1) >> 3));
                    cmax = max(cmax, c0 + 4 * k + ((31 - __clz((int)w[k])) >> 3));
                }
            rmin = min(rmin, row0 + r);
            rmax = max(rmax, row0 + r);
        }
    } else {
        for (size_t i = threadIdx.x; i < nbytes; i += kThreads)
            if (base[i]) {
                const int r = (int)(i / W), c = (int)(i - (size_t)r * W);
                rmin = min(rmin, row0 + r); rmax = max(rmax, row0 + r);
                cmin = min(cmin, c); cmax = max(cmax, c);
            }
    }
    rmin = __reduce_min_sync(0xffffffffu, rmin); rmax = __reduce_max_sync(0xffffffffu, rmax);
    cmin = __reduce_min_sync(0xffffffffu, cmin); cmax = __reduce_max_sync(0xffffffffu, cmax);
    if ((threadIdx.x & 31) == 0 && rmax >= 0) {
        atomicMin(&bounds[4 * b + 0], rmin); atomicMax(&bounds[4 * b + 1], rmax);
        atomicMin(&bounds[4 * b + 2], cmin); atomicMax(&bounds[4 * b + 3], cmax);
    }
}

__global__ void __launch_bounds__(kThreads)
k_roi_crop(const uint8_t* __restrict__ obj_bits, const uint8_t* __restrict__ hand_bits,
           const uint8_t* __restrict__ images_hwc, int H, int W, int S, float pad, float expansion,
           const int32_t* __restrict__ bounds, float* __restrict__ bbox, float* __restrict__ square_bbox,
           uint8_t* __restrict__ crop_mask, float* __restrict__ target, int8_t* __restrict__ target_tri,
           float* __restrict__ crop_image) {
    // run.py:49: (image / 255.).astype(float32) -- the 256 possible values, divided in double like numpy does
    __shared__ float s_lut[256];
    if (crop_image != nullptr) {
        for (int i = threadIdx.x; i < 256; i += kThreads) s_lut[i] = (float)((double)i / 255.0);
        __syncthreads();
    }
    const int b = blockIdx.y;
    const int idx = blockIdx.x * kThreads + threadIdx.x;
    const int r0 = bounds[4 * b + 0], r1 = bounds[4 * b + 1], c0 = bounds[4 * b + 2], c1 = bounds[4 * b + 3];
    const bool empty = r1 < 0;   // the reference's np.min raises on an empty mask: the host wrapper reports it
    float bb[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f}, xyxy[4] = {0.f, 0.f, 0.f, 0.f};
    if (!empty) roi_boxes(r0, r1, c0, c1, H, W, pad, expansion, bb, sq, xyxy);
    if (idx < 4) {
        bbox[4 * b + idx] = bb[idx];
        square_bbox[4 * b + idx] = sq[idx];
    }
    if (idx >= S * S) return;
    const size_t o = (size_t)b * S * S + idx;
    if (empty) {
        crop_mask[o] = 0;
        target[o] = 0.0f;
        if (target_tri) target_tri[o] = 0;
        if (crop_image)
            for (int c = 0; c < 3; c++) crop_image[((size_t)b * 3 + c) * S * S + idx] = 1.0f;
        return;
    }
    const int ph = idx / S, pw = idx - ph * S;
    const RoiGeom g = roi_geom(xyxy, S);
    const uint8_t* ob = obj_bits + (size_t)b * H * W;
    const bool obit = roi_align_cell(g, ph, pw, H, W, [&](int y, int x) { return (float)__ldg(ob + (size_t)y * W + x); }) >= 0.5f;
    bool hbit = false;
    if (hand_bits != nullptr && !obit) {   // the occluder only shows where the object is not
        const uint8_t* hb = hand_bits + (size_t)b * H * W;
        hbit = roi_align_cell(g, ph, pw, H, W, [&](int y, int x) { return (float)__ldg(hb + (size_t)y * W + x); }) >= 0.5f;
    }
    crop_mask[o] = obit ? 1 : 0;
    const float t = target_value(obit, hbit);
    target[o] = t;
    if (target_tri) target_tri[o] = (int8_t)t;
    if (crop_image != nullptr) {
        const uint8_t* im = images_hwc + (size_t)b * H * W * 3;
        float v[3] = {1.0f, 1.0f, 1.0f};   // run.py:51: the crop is white outside the object mask
        if (obit)
            roi_align_cell3(g, ph, pw, H, W, [&](int y, int x, float* px) {
                const uint8_t* q = im + ((size_t)y * W + x) * 3;
                px[0] = s_lut[__ldg(q)]; px[1] = s_lut[__ldg(q + 1)]; px[2] = s_lut[__ldg(q + 2)];
            }, v);
        for (int c = 0; c < 3; c++) crop_image[((size_t)b * 3 + c) * S * S + idx] = v[c];
    }
}

// The same crops for rendered TEMPLATE views (pose_initializtion.py:188-246, compute_prior_features): float RGBA
// renderings and a float depth map instead of a uint8 photograph, no occluder.  images: [B,H,W,pitch] floats (RGB in
// channels 0..2), depth: [B,H,W] or NULL.  crop_image [B,3,S,S] is white where the crop mask is 0 (:215), crop_depth
// [B,S,S] is the plain ROIAlign of the depth (:213-214).
__global__ void __launch_bounds__(kThreads)
k_roi_crop_f32(const uint8_t* __restrict__ obj_bits, const float* __restrict__ images, int pitch,
               const float* __restrict__ depth, int H, int W, int S, float pad, float expansion,
               const int32_t* __restrict__ bounds, float* __restrict__ bbox, float* __restrict__ square_bbox,
               uint8_t* __restrict__ crop_mask, float* __restrict__ crop_image, float* __restrict__ crop_depth) {
    const int b = blockIdx.y;
    const int idx = blockIdx.x * kThreads + threadIdx.x;
    const int r0 = bounds[4 * b + 0], r1 = bounds[4 * b + 1], c0 = bounds[4 * b + 2], c1 = bounds[4 * b + 3];
    const bool empty = r1 < 0;
    float bb[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f}, xyxy[4] = {0.f, 0.f, 0.f, 0.f};
    if (!empty) roi_boxes(r0, r1, c0, c1, H, W, pad, expansion, bb, sq, xyxy);
    if (idx < 4) {
        bbox[4 * b + idx] = bb[idx];
        square_bbox[4 * b + idx] = sq[idx];
    }
    if (idx >= S * S) return;
    const size_t o = (size_t)b * S * S + idx;
    if (empty) {
        crop_mask[o] = 0;
        for (int c = 0; c < 3; c++) crop_image[((size_t)b * 3 + c) * S * S + idx] = 1.0f;
        if (crop_depth) crop_depth[o] = 0.0f;
        return;
    }
    const int ph = idx / S, pw = idx - ph * S;
    const RoiGeom g = roi_geom(xyxy, S);
    const uint8_t* ob = obj_bits + (size_t)b * H * W;
    const bool obit = roi_align_cell(g, ph, pw, H, W, [&](int y, int x) { return (float)__ldg(ob + (size_t)y * W + x); }) >= 0.5f;
    crop_mask[o] = obit ? 1 : 0;
    float v[3] = {1.0f, 1.0f, 1.0f};
    if (obit) {
        const float* im = images + (size_t)b * H * W * pitch;
        roi_align_cell3(g, ph, pw, H, W, [&](int y, int x, float* px) {
            const float* q = im + ((size_t)y * W + x) * pitch;
            px[0] = __ldg(q); px[1] = __ldg(q + 1); px[2] = __ldg(q + 2);
        }, v);
    }
    for (int c = 0; c < 3; c++) crop_image[((size_t)b * 3 + c) * S * S + idx] = v[c];
    if (crop_depth != nullptr) {
        const float* dp = depth + (size_t)b * H * W;
        crop_depth[o] = roi_align_cell(g, ph, pw, H, W, [&](int y, int x) { return __ldg(dp + (size_t)y * W + x); });
    }
}

}  // namespace

extern "C" {

int dh_roi_process(const uint8_t* obj_bits, const uint8_t* hand_bits, const uint8_t* images_hwc, int32_t B,
                   int32_t H, int32_t W, int32_t S, float pad, float expansion, int32_t* bounds, float* bbox,
                   float* square_bbox, uint8_t* crop_mask, float* target, int8_t* target_tri, float* crop_image,
                   void* stream) {
    DH_REQUIRE(obj_bits && bounds && bbox && square_bbox && crop_mask && target, "NULL input / output");
    DH_REQUIRE(B > 0 && H > 0 && W > 0 && S > 0 && B <= 65535, "bad sizes");
    DH_REQUIRE((crop_image == nullptr) || (images_hwc != nullptr), "crop_image needs images");
    cudaStream_t st = (cudaStream_t)stream;
    k_roi_init<<<(4 * B + kThreads - 1) / kThreads, kThreads, 0, st>>>(bounds, B);
    DH_LAUNCH_OK("k_roi_init");
    k_roi_bounds<<<dim3((H + kRowsPerCta - 1) / kRowsPerCta, B), kThreads, 0, st>>>(obj_bits, H, W, bounds);
    DH_LAUNCH_OK("k_roi_bounds");
    k_roi_crop<<<dim3((S * S + kThreads - 1) / kThreads, B), kThreads, 0, st>>>(
        obj_bits, hand_bits, images_hwc, H, W, S, pad, expansion, bounds, bbox, square_bbox, crop_mask, target,
        target_tri, crop_image);
    DH_LAUNCH_OK("k_roi_crop");
    return DH_OK;
}

int dh_roi_process_f32(const uint8_t* obj_bits, const float* images, int32_t pitch, const float* depth, int32_t B,
                       int32_t H, int32_t W, int32_t S, float pad, float expansion, int32_t* bounds, float* bbox,
                       float* square_bbox, uint8_t* crop_mask, float* crop_image, float* crop_depth, void* stream) {
    DH_REQUIRE(obj_bits && images && bounds && bbox && square_bbox && crop_mask && crop_image, "NULL input / output");
    DH_REQUIRE(B > 0 && H > 0 && W > 0 && S > 0 && B <= 65535 && pitch >= 3, "bad sizes");
    DH_REQUIRE((crop_depth == nullptr) == (depth == nullptr), "depth and crop_depth go together");
    cudaStream_t st = (cudaStream_t)stream;
    k_roi_init<<<(4 * B + kThreads - 1) / kThreads, kThreads, 0, st>>>(bounds, B);
    DH_LAUNCH_OK("k_roi_init");
    k_roi_bounds<<<dim3((H + kRowsPerCta - 1) / kRowsPerCta, B), kThreads, 0, st>>>(obj_bits, H, W, bounds);
    DH_LAUNCH_OK("k_roi_bounds");
    k_roi_crop_f32<<<dim3((S * S + kThreads - 1) / kThreads, B), kThreads, 0, st>>>(
        obj_bits, images, pitch, depth, H, W, S, pad, expansion, bounds, bbox, square_bbox, crop_mask, crop_image,
        crop_depth);
    DH_LAUNCH_OK("k_roi_crop_f32");
    return DH_OK;
}

}  // extern "C"
