// dh_roi_core.h -- arithmetic of the ROI preprocessing (run.py:26-72 process_input), shared by the CUDA kernels
// (dh_roi.cu) and the host emulation of the CPU tests (tests/emu).  fp32 with no contraction (--fmad=false /
// -ffp-contract=off): every expression below keeps the operation order of the code it restates, so that the
// thresholded crops come out bit-identical.
//
//   boxes      run.py:37-46, utils/bbox.py:73-117 (make_bbox_square, BoxMode XYXY <-> XYWH), numpy / torch float32
//   ROIAlign   detectron2 v0.4 ROIAlign((S, S), 1.0, 0, aligned=True) == torchvision.ops.roi_align, CPU kernel
//              torchvision/csrc/ops/cpu/roi_align_kernel.cpp (roi_align_forward_kernel_impl +
//              pre_calc_for_bilinear_interpolate), T = float
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef DH_HD
#if defined(__CUDACC__)
#define DH_HD __host__ __device__ __forceinline__
#else
#define DH_HD inline
#endif
#endif

namespace dh {

// Tight pixel bounds of the mask -> bbox (x, y, w, h), the square box around it (x, y, b, b) and its corners.
// run.py:37-46: rows/cols padded by `pad` and clamped to the image, bbox_xy_to_wh, make_bbox_square(bbox, expansion).
DH_HD void roi_boxes(int min_row, int max_row, int min_col, int max_col, int H, int W, float pad, float expansion,
                     float* bbox_xywh, float* square_xywh, float* square_xyxy) {
    const float y0 = fmaxf((float)min_row - pad, 0.0f), y1 = fminf((float)max_row + pad, (float)H);
    const float x0 = fmaxf((float)min_col - pad, 0.0f), x1 = fminf((float)max_col + pad, (float)W);
    const float w = x1 - x0, h = y1 - y0;
    bbox_xywh[0] = x0; bbox_xywh[1] = y0; bbox_xywh[2] = w; bbox_xywh[3] = h;
    const float cx = x0 + w / 2.0f, cy = y0 + h / 2.0f;
    float b = fmaxf(w, h);
    b = b * (1.0f + expansion);               // numpy: float32 array *= python float -> float32 product
    const float sx = cx - b / 2.0f, sy = cy - b / 2.0f;
    square_xywh[0] = sx; square_xywh[1] = sy; square_xywh[2] = b; square_xywh[3] = b;
    square_xyxy[0] = sx; square_xyxy[1] = sy; square_xyxy[2] = b + sx; square_xyxy[3] = b + sy;
}

struct RoiGeom {
    float start_w, start_h, bin_w, bin_h, count;
    int grid_w, grid_h;
};

// roi_align_forward_kernel_impl: spatial_scale = 1, aligned = true, sampling_ratio = 0 (adaptive grid)
DH_HD RoiGeom roi_geom(const float* xyxy, int pooled) {
    RoiGeom g;
    const float offset = 0.5f;
    g.start_w = xyxy[0] * 1.0f - offset;
    g.start_h = xyxy[1] * 1.0f - offset;
    const float end_w = xyxy[2] * 1.0f - offset, end_h = xyxy[3] * 1.0f - offset;
    const float roi_width = end_w - g.start_w, roi_height = end_h - g.start_h;
    g.bin_h = roi_height / (float)pooled;
    g.bin_w = roi_width / (float)pooled;
    g.grid_h = (int)ceilf(roi_height / (float)pooled);
    g.grid_w = (int)ceilf(roi_width / (float)pooled);
    const int n = g.grid_h * g.grid_w;
    g.count = (float)(n > 1 ? n : 1);
    return g;
}

// One output cell (ph, pw) of one channel: the adaptive grid of bilinear samples, summed in (iy, ix) order, over
// the count.  fetch(y, x) returns the channel's value at an integer pixel as float.
template <typename Fetch>
DH_HD float roi_align_cell(const RoiGeom& g, int ph, int pw, int height, int width, Fetch fetch) {
    float out = 0.0f;
    for (int iy = 0; iy < g.grid_h; iy++) {
        const float yy = g.start_h + (float)ph * g.bin_h + ((float)iy + 0.5f) * g.bin_h / (float)g.grid_h;
        for (int ix = 0; ix < g.grid_w; ix++) {
            const float xx = g.start_w + (float)pw * g.bin_w + ((float)ix + 0.5f) * g.bin_w / (float)g.grid_w;
            float x = xx, y = yy;
            if (y < -1.0f || y > (float)height || x < -1.0f || x > (float)width) continue;  // four zero weights
            if (y <= 0.0f) y = 0.0f;
            if (x <= 0.0f) x = 0.0f;
            int y_low = (int)y, x_low = (int)x, y_high, x_high;
            if (y_low >= height - 1) { y_high = y_low = height - 1; y = (float)y_low; } else { y_high = y_low + 1; }
            if (x_low >= width - 1) { x_high = x_low = width - 1; x = (float)x_low; } else { x_high = x_low + 1; }
            const float ly = y - (float)y_low, lx = x - (float)x_low;
            const float hy = 1.0f - ly, hx = 1.0f - lx;
            const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
            float t = w1 * fetch(y_low, x_low);
            t = t + w2 * fetch(y_low, x_high);
            t = t + w3 * fetch(y_high, x_low);
            t = t + w4 * fetch(y_high, x_high);
            out = out + t;
        }
    }
    return out / g.count;
}

// The same cell for three channels that share the sample positions and weights (an H x W x 3 image): every
// channel's sum runs in the same order as roi_align_cell's, so the results are identical to three separate calls.
// fetch3(y, x, v) fills v[0..2].
template <typename Fetch3>
DH_HD void roi_align_cell3(const RoiGeom& g, int ph, int pw, int height, int width, Fetch3 fetch3, float* out3) {
    float out[3] = {0.0f, 0.0f, 0.0f};
    for (int iy = 0; iy < g.grid_h; iy++) {
        const float yy = g.start_h + (float)ph * g.bin_h + ((float)iy + 0.5f) * g.bin_h / (float)g.grid_h;
        for (int ix = 0; ix < g.grid_w; ix++) {
            const float xx = g.start_w + (float)pw * g.bin_w + ((float)ix + 0.5f) * g.bin_w / (float)g.grid_w;
            float x = xx, y = yy;
            if (y < -1.0f || y > (float)height || x < -1.0f || x > (float)width) continue;
            if (y <= 0.0f) y = 0.0f;
            if (x <= 0.0f) x = 0.0f;
            int y_low = (int)y, x_low = (int)x, y_high, x_high;
            if (y_low >= height - 1) { y_high = y_low = height - 1; y = (float)y_low; } else { y_high = y_low + 1; }
            if (x_low >= width - 1) { x_high = x_low = width - 1; x = (float)x_low; } else { x_high = x_low + 1; }
            const float ly = y - (float)y_low, lx = x - (float)x_low;
            const float hy = 1.0f - ly, hx = 1.0f - lx;
            const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
            float v1[3], v2[3], v3[3], v4[3];
            fetch3(y_low, x_low, v1); fetch3(y_low, x_high, v2); fetch3(y_high, x_low, v3); fetch3(y_high, x_high, v4);
            for (int c = 0; c < 3; c++) {
                float t = w1 * v1[c];
                t = t + w2 * v2[c];
                t = t + w3 * v3[c];
                t = t + w4 * v4[c];
                out[c] = out[c] + t;
            }
        }
    }
    for (int c = 0; c < 3; c++) out3[c] = out[c] / g.count;
}

// utils/maskutils.py:8-30 for one object and one occluder layer: 1 object, -1 occluder where it does not cover the
// object, 0 background
DH_HD float target_value(bool object_bit, bool occluder_bit) { return object_bit ? 1.0f : (occluder_bit ? -1.0f : 0.0f); }

}  // namespace dh
