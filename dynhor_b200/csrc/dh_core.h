// dh_core.h -- per-pose / per-vertex / per-face / per-pixel arithmetic of the joint-optimisation hot path.
//
// Shared by the sm_100a kernels (dh_jointopt.cu) and by tests/emu/ (a host build of the same functions, used
// only by the CPU test-suite to check the kernel logic against the oracle without a GPU; it is NOT a product
// path and the Python package never loads it).
//
// fp32 contract: every function below performs plain IEEE-754 single precision operations in a fixed order.
// The translation units that include this header are compiled with --fmad=false (nvcc) / -ffp-contract=off
// (g++): an FMA exists only where fmaf() is written, which is where torch's CPU matmul (the oracle) uses one.
//
// Reference behaviour restated here (file:line under /root/reference/ObjTracker):
//   rot6d_to_R           utils/geometry.py:19-25
//   transform_vertex     utils/camera.py:204-206        ((|s|*v) @ R + T)
//   project_vertex       utils/camera.py:39-62          (K in unit-image coordinates, orig_size, eps 1e-9)
//   face/pixel/backward  third-party neural_renderer kernels called at utils/losses.py:68
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DH_HD __host__ __device__ __forceinline__
#else
#define DH_HD inline
#endif

namespace dh {

// ---------------------------------------------------------------------------------------------- helpers
DH_HD int f2i_sat(float v) {  // CUDA float->int conversion semantics: truncate, saturate, NaN -> 0
#if defined(__CUDA_ARCH__)
    return __float2int_rz(v);
#else
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (-2147483647 - 1);
    return (int)v;
#endif
}
DH_HD int ctz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
DH_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
DH_HD bool finite3(float a, float b, float c) {
    return (fabsf(a) <= 3.0e38f) && (fabsf(b) <= 3.0e38f) && (fabsf(c) <= 3.0e38f);
}

// ---------------------------------------------------------------------------------------------- pose
// rot6d: [3][2] row-major (columns a1, a2).  R: [3][3] row-major whose COLUMNS are b1, b2, b3.
DH_HD void rot6d_to_R(const float* r6, float* R) {
    const float a1[3] = {r6[0], r6[2], r6[4]};
    const float a2[3] = {r6[1], r6[3], r6[5]};
    float n1 = sqrtf((a1[0] * a1[0] + a1[1] * a1[1]) + a1[2] * a1[2]);
    n1 = fmaxf(n1, 1e-12f);
    const float b1[3] = {a1[0] / n1, a1[1] / n1, a1[2] / n1};
    const float d = (b1[0] * a2[0] + b1[1] * a2[1]) + b1[2] * a2[2];
    const float u2[3] = {a2[0] - d * b1[0], a2[1] - d * b1[1], a2[2] - d * b1[2]};
    // torch's CPU kernels (the oracle): the norm of the contiguous u2 is an FMA chain, the norm of the strided
    // a1 above is not; cross is fma(a, b, -(c * d)).  Mirrored so that R is bit-identical to the oracle's.
    float n2 = sqrtf(fmaf(u2[2], u2[2], fmaf(u2[1], u2[1], u2[0] * u2[0])));
    n2 = fmaxf(n2, 1e-12f);
    const float b2[3] = {u2[0] / n2, u2[1] / n2, u2[2] / n2};
    const float b3[3] = {fmaf(b1[1], b2[2], -(b1[2] * b2[1])), fmaf(b1[2], b2[0], -(b1[0] * b2[2])),
                         fmaf(b1[0], b2[1], -(b1[1] * b2[0]))};
    for (int i = 0; i < 3; i++) {
        R[3 * i + 0] = b1[i];
        R[3 * i + 1] = b2[i];
        R[3 * i + 2] = b3[i];
    }
}

// Backward of rot6d_to_R: G = dL/dR ([3][3] row-major) -> g6 = dL/drot6d ([3][2] row-major).  double math.
DH_HD void rot6d_backward(const float* r6, const double* G, double* g6) {
    const double a1[3] = {r6[0], r6[2], r6[4]};
    const double a2[3] = {r6[1], r6[3], r6[5]};
    double n1 = sqrt(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]);
    const bool c1 = n1 < 1e-12;
    if (c1) n1 = 1e-12;
    const double b1[3] = {a1[0] / n1, a1[1] / n1, a1[2] / n1};
    const double d = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
    const double u2[3] = {a2[0] - d * b1[0], a2[1] - d * b1[1], a2[2] - d * b1[2]};
    double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    const bool c2 = n2 < 1e-12;
    if (c2) n2 = 1e-12;
    const double b2[3] = {u2[0] / n2, u2[1] / n2, u2[2] / n2};
    double g1[3] = {G[0], G[3], G[6]}, g2[3] = {G[1], G[4], G[7]};
    const double g3[3] = {G[2], G[5], G[8]};
    // b3 = b1 x b2:  dL/db1 += b2 x g3,  dL/db2 += g3 x b1
    g1[0] += b2[1] * g3[2] - b2[2] * g3[1];
    g1[1] += b2[2] * g3[0] - b2[0] * g3[2];
    g1[2] += b2[0] * g3[1] - b2[1] * g3[0];
    g2[0] += g3[1] * b1[2] - g3[2] * b1[1];
    g2[1] += g3[2] * b1[0] - g3[0] * b1[2];
    g2[2] += g3[0] * b1[1] - g3[1] * b1[0];
    // b2 = u2 / n2
    double gu[3];
    if (c2) {
        for (int i = 0; i < 3; i++) gu[i] = g2[i] / n2;
    } else {
        const double s = b2[0] * g2[0] + b2[1] * g2[1] + b2[2] * g2[2];
        for (int i = 0; i < 3; i++) gu[i] = (g2[i] - s * b2[i]) / n2;
    }
    // u2 = a2 - d b1, d = b1.a2
    const double gd = -(gu[0] * b1[0] + gu[1] * b1[1] + gu[2] * b1[2]);
    double ga2[3];
    for (int i = 0; i < 3; i++) {
        ga2[i] = gu[i] + gd * b1[i];
        g1[i] += -d * gu[i] + gd * a2[i];
    }
    // b1 = a1 / n1
    double ga1[3];
    if (c1) {
        for (int i = 0; i < 3; i++) ga1[i] = g1[i] / n1;
    } else {
        const double s = b1[0] * g1[0] + b1[1] * g1[1] + b1[2] * g1[2];
        for (int i = 0; i < 3; i++) ga1[i] = (g1[i] - s * b1[i]) / n1;
    }
    for (int i = 0; i < 3; i++) {
        g6[2 * i + 0] = ga1[i];
        g6[2 * i + 1] = ga2[i];
    }
}

// (|s| v) @ R + T, the k-ordered FMA chain of a K=3 matmul.  Matches torch's CPU bmm (the oracle) bit for bit
// for meshes of more than 44 vertices (below that torch takes a non-FMA small-matrix path).
DH_HD void transform_vertex(const float* v, float s_abs, const float* R, const float* T, float* out) {
    const float s0 = s_abs * v[0], s1 = s_abs * v[1], s2 = s_abs * v[2];
    for (int j = 0; j < 3; j++) {
        const float acc = fmaf(s2, R[6 + j], fmaf(s1, R[3 + j], s0 * R[j]));
        out[j] = acc + T[j];
    }
}

// camera-space (x,y,z) -> rasteriser NDC (u,v) + depth z.  K: [3][3] row-major (third row unused).
DH_HD void project_vertex(const float* c, const float* K, float orig, float* u_out, float* v_out) {
    const float zc = c[2] + 1e-9f;
    const float x_ = c[0] / zc, y_ = c[1] / zc;
    float u = fmaf(1.0f, K[2], fmaf(y_, K[1], x_ * K[0]));
    float v = fmaf(1.0f, K[5], fmaf(y_, K[4], x_ * K[3]));
    v = orig - v;
    const float half = orig / 2.0f;
    *u_out = (2.0f * (u - half)) / orig;
    *v_out = (2.0f * (v - half)) / orig;
}

// Backward of project_vertex: (gu, gv) on NDC u,v -> gradient on camera-space (x,y,z).
DH_HD void project_vertex_backward(const float* c, const float* K, float orig, float gu, float gv, float* g) {
    const float zc = c[2] + 1e-9f;
    const float x_ = c[0] / zc, y_ = c[1] / zc;
    const float gu1 = (gu * 2.0f) / orig;
    const float gv1 = -((gv * 2.0f) / orig);
    const float gx_ = gu1 * K[0] + gv1 * K[3];
    const float gy_ = gu1 * K[1] + gv1 * K[4];
    g[0] = gx_ / zc;
    g[1] = gy_ / zc;
    g[2] = -(gx_ * x_ + gy_ * y_) / zc;
}

// ---------------------------------------------------------------------------------------------- raster forward
DH_HD float ndc_to_pix(float v, int is) {  // [-1,1] -> [0, is-1]
    float t = v * (float)is;
    t = t + (float)is;
    t = t - 1.0f;
    return 0.5f * t;
}
DH_HD float pix_to_ndc(int i, int is) { return (float)(2 * i + 1 - is) / (float)is; }

DH_HD bool face_backside(float x0, float y0, float x1, float y1, float x2, float y2) {
    return ((y2 - y0) * (x1 - x0)) < ((y1 - y0) * (x2 - x0));
}

struct FaceSetup {
    float x[3], y[3], z[3];  // NDC x,y and depth of the three vertices in this winding's order
    float inv[9];            // inverse of [[px0,px1,px2],[py0,py1,py2],[1,1,1]] (pixel coordinates)
    int x_lo, x_hi, y_lo, y_hi;  // conservative pixel bounding box, clamped to the image (empty if culled)
};

#define kBoxSlack 0.015625f

#if defined(__CUDA_ARCH__)
#define DH_APPROX_DIV(a, b) __fdividef((a), (b))
#else
#define DH_APPROX_DIV(a, b) ((a) / (b))
#endif

// Conservative pixel bbox of a face; false if the face cannot produce a recorded pixel (back side, non-finite
// coordinates or fully off-screen).
DH_HD bool face_bbox(const float* x, const float* y, int is, int* x_lo, int* x_hi, int* y_lo, int* y_hi) {
    if (!finite3(x[0], x[1], x[2]) || !finite3(y[0], y[1], y[2])) return false;
    if (face_backside(x[0], y[0], x[1], y[1], x[2], y[2])) return false;
    const float px0 = ndc_to_pix(x[0], is), px1 = ndc_to_pix(x[1], is), px2 = ndc_to_pix(x[2], is);
    const float py0 = ndc_to_pix(y[0], is), py1 = ndc_to_pix(y[1], is), py2 = ndc_to_pix(y[2], is);
    const float lim = (float)(is + 2);
    const float xmin = fminf(fmaxf(fminf(px0, fminf(px1, px2)), -3.0f), lim);
    const float xmax = fminf(fmaxf(fmaxf(px0, fmaxf(px1, px2)), -3.0f), lim);
    const float ymin = fminf(fmaxf(fminf(py0, fminf(py1, py2)), -3.0f), lim);
    const float ymax = fminf(fmaxf(fmaxf(py0, fmaxf(py1, py2)), -3.0f), lim);
    // A pixel centre can pass the three fp32 edge tests only within ~1e-6 px of the exact triangle (rounding of
    // the edge functions near the face), so 1/64 px of slack around the vertices' bounding box is conservative.
    int xl = (int)ceilf(xmin - kBoxSlack), xh = (int)floorf(xmax + kBoxSlack);
    int yl = (int)ceilf(ymin - kBoxSlack), yh = (int)floorf(ymax + kBoxSlack);
    if (xl < 0) xl = 0;
    if (yl < 0) yl = 0;
    if (xh > is - 1) xh = is - 1;
    if (yh > is - 1) yh = is - 1;
    *x_lo = xl; *x_hi = xh; *y_lo = yl; *y_hi = yh;
    return xl <= xh && yl <= yh;
}

DH_HD void face_inverse(FaceSetup& f, int is) {
    float p[3][2];
    for (int k = 0; k < 3; k++) {
        p[k][0] = ndc_to_pix(f.x[k], is);
        p[k][1] = ndc_to_pix(f.y[k], is);
    }
    float fi[9];
    fi[0] = p[1][1] - p[2][1];
    fi[1] = p[2][0] - p[1][0];
    fi[2] = p[1][0] * p[2][1] - p[2][0] * p[1][1];
    fi[3] = p[2][1] - p[0][1];
    fi[4] = p[0][0] - p[2][0];
    fi[5] = p[2][0] * p[0][1] - p[0][0] * p[2][1];
    fi[6] = p[0][1] - p[1][1];
    fi[7] = p[1][0] - p[0][0];
    fi[8] = p[0][0] * p[1][1] - p[1][0] * p[0][1];
    float den = p[2][0] * (p[0][1] - p[1][1]);
    den = den + p[0][0] * (p[1][1] - p[2][1]);
    den = den + p[1][0] * (p[2][1] - p[0][1]);
    for (int k = 0; k < 9; k++) f.inv[k] = fi[k] / den;
}

// Conservative pixel interval [xa, xb] of row `yp` (NDC) that can pass the three edge tests, clipped to
// [x_lo, x_hi]: every edge with dy != 0 bounds x from one side.  Only a pre-filter -- pixel_inside still decides --
// so approximate division plus 1/64 px of slack is enough.
DH_HD void row_span(const FaceSetup& f, float yp, int is, int x_lo, int x_hi, int* xa, int* xb) {
    float lo = -3.0e38f, hi = 3.0e38f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int j = (i + 1) % 3;
        const float dx = f.x[j] - f.x[i], dy = f.y[j] - f.y[i];
        if (dy != 0.0f) {
            const float xb_ = f.x[i] + DH_APPROX_DIV((yp - f.y[i]) * dx, dy);
            if (dy > 0.0f) hi = fminf(hi, xb_);
            else           lo = fmaxf(lo, xb_);
        }
    }
    // NDC -> pixel coordinates, clamped so the int conversion is safe
    const float plo = fminf(fmaxf(0.5f * (lo * (float)is + (float)is - 1.0f), -2.0f), (float)(is + 2));
    const float phi = fminf(fmaxf(0.5f * (hi * (float)is + (float)is - 1.0f), -2.0f), (float)(is + 2));
    const int a = (int)ceilf(plo - kBoxSlack), b = (int)floorf(phi + kBoxSlack);
    *xa = a > x_lo ? a : x_lo;
    *xb = b < x_hi ? b : x_hi;
}

// true if pixel centre (xp,yp) [NDC] is not rejected by any of the three edge tests
DH_HD bool pixel_inside(const FaceSetup& f, float xp, float yp) {
    if ((yp - f.y[0]) * (f.x[1] - f.x[0]) < (xp - f.x[0]) * (f.y[1] - f.y[0])) return false;
    if ((yp - f.y[1]) * (f.x[2] - f.x[1]) < (xp - f.x[1]) * (f.y[2] - f.y[1])) return false;
    if ((yp - f.y[2]) * (f.x[0] - f.x[2]) < (xp - f.x[2]) * (f.y[0] - f.y[2])) return false;
    return true;
}

// depth of the face at pixel (xi,yi); returns false if rejected by the near/far test (or NaN)
DH_HD bool pixel_depth(const FaceSetup& f, int xi, int yi, float near, float far, float* zp_out) {
    float w[3];
    for (int k = 0; k < 3; k++) {
        float t = f.inv[3 * k + 0] * (float)xi;
        t = t + f.inv[3 * k + 1] * (float)yi;
        w[k] = t + f.inv[3 * k + 2];
    }
    float w_sum = 0.0f;
    for (int k = 0; k < 3; k++) {
        w[k] = fminf(fmaxf(w[k], 0.0f), 1.0f);
        w_sum = w_sum + w[k];
    }
    for (int k = 0; k < 3; k++) w[k] = w[k] / w_sum;
    float s = w[0] / f.z[0];
    s = s + w[1] / f.z[1];
    s = s + w[2] / f.z[2];
    const float zp = 1.0f / s;
    if (zp <= near || far <= zp) return false;
    if (!(zp < far)) return false;  // NaN never beats the initial depth
    *zp_out = zp;
    return true;
}

// z-buffer key: depth bits in the high word, face number in the low word.  Depths are positive, so the
// unsigned order of the key is (depth, face number): min() == nearest, lowest face number on exact ties.
DH_HD unsigned long long zkey(float zp, int fn) {
#if defined(__CUDA_ARCH__)
    return ((unsigned long long)__float_as_uint(zp) << 32) | (unsigned int)fn;
#else
    union { float f; uint32_t u; } c;
    c.f = zp;
    return ((unsigned long long)c.u << 32) | (unsigned int)fn;
#endif
}
#define DH_ZKEY_EMPTY 0xFFFFFFFFFFFFFFFFull

// ---------------------------------------------------------------------------------------------- raster backward
// Maps a face needs for the edge-scan pseudo-gradient.  Bitmaps are 32 pixels per word, bit i = pixel 32*w+i.
struct BwdMaps {
    const uint32_t* alpha;     // [is][is/32]  row-major coverage (rasteriser row order, i.e. before the flip)
    const uint32_t* neg;       // [is][is/32]  row-major: alpha == 0 && grad < 0   (optional: NULL -> derived from
                               //              alpha and neg_pool on the fly)
    const uint32_t* negT;      // [is][is/32]  column-major copy of `neg`: negT[c][r/32]
    const uint32_t* pos_pool;  // [S][ceil(S/32)] output-resolution bitmap: grad > 0
    const uint32_t* neg_pool;  // [S][ceil(S/32)] output-resolution bitmap: grad < 0
    const int16_t* row_lo;     // [is] first / last set bit of every row / column of `neg` (lo > hi: none);
    const int16_t* row_hi;     //      optional (NULL -> no range filter)
    const int16_t* col_lo;
    const int16_t* col_hi;
    const float* gpool;        // [S][S]  dL/d(rendered silhouette) at output resolution
    const int32_t* fidx;       // [is][is] face index map (-1 none)
    int is, S, aa, wpr, wpr_pool;
    float gscale;  // 0.25 with anti-aliasing (average-pool backward), else 1
};
DH_HD int cell_y(const BwdMaps& m, int r) { const int rf = m.is - 1 - r; return m.aa ? (rf >> 1) : rf; }
DH_HD int cell_x(const BwdMaps& m, int c) { return m.aa ? (c >> 1) : c; }
DH_HD bool alpha_at(const BwdMaps& m, int r, int c) { return (m.alpha[r * m.wpr + (c >> 5)] >> (c & 31)) & 1u; }
DH_HD float grad_at(const BwdMaps& m, int r, int c) { return m.gpool[cell_y(m, r) * m.S + cell_x(m, c)] * m.gscale; }
DH_HD bool pos_at(const BwdMaps& m, int r, int c) {
    const int x = cell_x(m, c);
    return (m.pos_pool[cell_y(m, r) * m.wpr_pool + (x >> 5)] >> (x & 31)) & 1u;
}
DH_HD uint32_t spread16(uint32_t x) {  // bit i -> bits 2i and 2i+1
    x &= 0xFFFFu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x | (x << 1);
}
// word w of row r of the "uncovered and gradient < 0" bitmap, from the coverage and the pooled sign bitmap
DH_HD uint32_t neg_row_word(const uint32_t* alpha, const uint32_t* neg_pool, int is, int aa, int wpr, int wprp, int r,
                            int w) {
    const int rf = is - 1 - r;
    uint32_t nb;
    if (aa) {
        const uint32_t pw = neg_pool[(rf >> 1) * wprp + (w >> 1)];
        nb = spread16((w & 1) ? (pw >> 16) : pw);
    } else {
        nb = neg_pool[rf * wprp + w];
    }
    return ~alpha[r * wpr + w] & nb;
}
// word w of the scan line of (axis, d0): column d0 (axis 0) or row d0 (axis 1)
DH_HD uint32_t neg_line_word(const BwdMaps& m, int axis, int d0, int w) {
    if (axis == 0) return m.negT[d0 * m.wpr + w];
    if (m.neg != nullptr) return m.neg[d0 * m.wpr + w];
    return neg_row_word(m.alpha, m.neg_pool, m.is, m.aa, m.wpr, m.wpr_pool, d0, w);
}

// ---- span list of a backward batch: the non-empty (edge, axis) spans of the batch's <= 32 faces, compacted in (lane,
// span) order; a span's crossings (scan lines d0_from ... d0_to) are numbered start ... start + length - 1, and the
// batch's crossings 0 ... T-1 are handed to the lanes 32 at a time.
//   info    = lane | (edge * 2 + axis) << 5 | (direction > 0) << 8 | (d0_from - start + 2^17) << 9
//   start16 = start mod 2^16 (T can pass 2^16 -- 32 faces across a 512-pixel raster: 65 568, a triangle's spans add
//             up to ~2 (width + height) scan lines; only differences below 2^14 are ever formed)
DH_HD uint32_t span_info(int lane, int k, int d0_from, bool dpos, uint32_t start) {
    return ((uint32_t)lane | ((uint32_t)k << 5) | (dpos ? (1u << 8) : 0u) | ((uint32_t)(d0_from + (1 << 17)) << 9)) -
           (start << 9);
}
DH_HD int span_info_d0(uint32_t info, int idx) { return idx + (int)(info >> 9) - (1 << 17); }
// Step `base` (crossings base ... base + 31): every span after s0 -- s0 starts at or before `base` -- that starts inside
// the step marks the bit of its first crossing; at most 32 spans can (each has a crossing), so the 32 lanes looking at
// spans s0 + 1 ... s0 + 32 see them all.  With M = OR of the marks, lane l works on span s0 + popc(M & bits 0..l), and
// the next step's s0 is s0 + popc(M).
DH_HD uint32_t span_mark(uint16_t start16, int base) {
    const uint32_t rel = (uint32_t)(uint16_t)(start16 - (uint16_t)base);
    return rel < 32u ? (1u << rel) : 0u;
}
DH_HD int span_of_lane(int s0, uint32_t M, int lane) { return s0 + popc32(M & (0xFFFFFFFFu >> (31 - lane))); }

// Guided batch schedule of the backward's items (a function of the item count only, so that the sums of a chunk are
// added up in the same order whatever warp ran which batch): full 32-item batches while more than a round of them is
// left, then remaining / (warps * div), never fewer than `min_items`; no crumbs at the end.  starts[0 .. nb] (starts[nb]
// = n_items); at most max_batches - 1 batches.  Returns nb.
DH_HD int bwd_guided_schedule(int n_items, int warps, int min_items, int div, int max_batches, uint16_t* starts) {
    int pos = 0, nb = 0;
    while (pos < n_items) {
        const int rem = n_items - pos;
        int sz = rem / (warps * div);
        if (sz < min_items) sz = min_items;
        if (sz > 32) sz = 32;
        if (rem - sz < min_items / 2) sz = rem < 32 ? rem : 32;
        if (nb + (rem + 31) / 32 >= max_batches - 1) sz = rem < 32 ? rem : 32;   // (never with <= 2048 items)
        starts[nb++] = (uint16_t)pos;
        pos += sz;
    }
    starts[nb] = (uint16_t)n_items;
    return nb;
}

// One edge of a face seen along one axis: p0 -> p1 is the edge, p2 the opposite vertex; component 0 is the
// coordinate the scan steps along (d0), component 1 the one it scans (d1).
struct Span {
    float p00, p01, p10, p11, p20, p21, slope;
    int direction, d0_from, d0_to;
};
// px, py: pixel coordinates of the three vertices (ndc_to_pix), in the winding's order.
DH_HD void span_setup(const float* px, const float* py, int edge, int axis, int is, Span& sp) {
    const int i0 = edge, i1 = (edge + 1) % 3, i2 = (edge + 2) % 3;
    sp.p00 = axis ? py[i0] : px[i0]; sp.p01 = axis ? px[i0] : py[i0];
    sp.p10 = axis ? py[i1] : px[i1]; sp.p11 = axis ? px[i1] : py[i1];
    sp.p20 = axis ? py[i2] : px[i2]; sp.p21 = axis ? px[i2] : py[i2];
    if (axis == 0) sp.direction = (sp.p00 < sp.p10) ? -1 : 1;
    else           sp.direction = (sp.p00 < sp.p10) ? 1 : -1;
    sp.d0_from = f2i_sat(fmaxf(ceilf(fminf(sp.p00, sp.p10)), 0.0f));
    sp.d0_to = f2i_sat(fminf(fmaxf(sp.p00, sp.p10), (float)(is - 1)));
    sp.slope = (sp.p11 - sp.p01) / (sp.p10 - sp.p00);
}
// Crossing of the edge with scan line d0: the pixel just inside (d1_in) and just outside (d1_out) the face.
DH_HD bool span_crossing(const Span& sp, int d0, int is, float* d1_cross, int* d1_in, int* d1_out) {
    float c = sp.slope * ((float)d0 - sp.p00);
    c = c + sp.p01;
    int in;
    if (0 < sp.direction) in = f2i_sat(floorf(c));
    else                  in = f2i_sat(ceilf(c));
    const int out = in + sp.direction;
    *d1_cross = c; *d1_in = in; *d1_out = out;
    if (in < 0 || is <= in) return false;
    if (out < 0 || is <= out) return false;
    return true;
}
DH_HD void out_scan_range(int direction, int d1_out, int is, int* from, int* to) {
    const int d1_limit = (0 < direction) ? is - 1 : 0;
    int f = d1_out < d1_limit ? d1_out : d1_limit;
    if (f < 0) f = 0;
    int t = d1_out > d1_limit ? d1_out : d1_limit;
    if (t > is - 1) t = is - 1;
    *from = f; *to = t;
}
DH_HD void in_scan_range(const Span& sp, int d0, int d1_in, int is, int* from, int* to) {
    float c2;
    if (((float)d0 - sp.p00) * ((float)d0 - sp.p20) < 0.0f) {
        c2 = (sp.p21 - sp.p01) / (sp.p20 - sp.p00);
        c2 = c2 * ((float)d0 - sp.p00);
        c2 = c2 + sp.p01;
    } else {
        c2 = (sp.p11 - sp.p21) / (sp.p10 - sp.p20);
        c2 = c2 * ((float)d0 - sp.p20);
        c2 = c2 + sp.p21;
    }
    int d1_limit;
    if (0 < sp.direction) d1_limit = f2i_sat(ceilf(c2));
    else                  d1_limit = f2i_sat(floorf(c2));
    int f = d1_in < d1_limit ? d1_in : d1_limit;
    if (f < 0) f = 0;
    int t = d1_in > d1_limit ? d1_in : d1_limit;
    if (t > is - 1) t = is - 1;
    *from = f; *to = t;
}
// The two terms (diff / dist) one pixel d1 of the scan contributes to the edge's vertices p0 (ta) and p1 (tb);
// the gradient accumulates MINUS these.
DH_HD void edge_terms(float diff, int d0, int d1, float d1_cross, float p00, float p10, float eps, int is, float* ta,
                      float* tb) {
    *ta = 0.0f; *tb = 0.0f;
    if (p10 != (float)d0) {
        float t = (p10 - p00) / (p10 - (float)d0);
        t = t * ((float)d1 - d1_cross);
        float dist = (t * 2.0f) / (float)is;
        dist = (0.0f < dist) ? dist + eps : dist - eps;
        *ta = diff / dist;
    }
    if (p00 != (float)d0) {
        float t = (p10 - p00) / ((float)d0 - p00);
        t = t * ((float)d1 - d1_cross);
        float dist = (t * 2.0f) / (float)is;
        dist = (0.0f < dist) ? dist + eps : dist - eps;
        *tb = diff / dist;
    }
}

// The same two terms with the per-crossing factors hoisted out of the pixel loop (kernel path).  The final
// quotient uses the approximate divider on the device: gradients carry a 1e-3 bar, not a bit-exact one.
struct EdgeCoef { float ka, kb; };  // 0 where the reference skips the term (edge end point exactly on the line)
DH_HD void edge_coefs(float p00, float p10, int d0, EdgeCoef& c) {
    c.ka = (p10 != (float)d0) ? (p10 - p00) / (p10 - (float)d0) : 0.0f;
    c.kb = (p00 != (float)d0) ? (p10 - p00) / ((float)d0 - p00) : 0.0f;
}
// Branch-free: k == 0 (term skipped by the reference) yields 0.  dist = k (d1 - d1_cross) (2/is) +- eps never gets
// below eps in magnitude, so the raw reciprocal approximation is safe.
DH_HD float edge_term_fast(float k, float diff, int d1, float d1_cross, float eps, float two_over_is) {
    float dist = (k * ((float)d1 - d1_cross)) * two_over_is;
    dist = (0.0f < dist) ? dist + eps : dist - eps;
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(dist));
    const float t = diff * r;
#else
    const float t = diff / dist;
#endif
    return (k != 0.0f) ? t : 0.0f;
}

// Pseudo-gradient of the loss w.r.t. the NDC (x,y) of the three vertices of face `fn`, one face per call (the
// serial form: used by the host emulation and as the definition the warp-cooperative kernel must reproduce).
// grad: [3][2] (vertex, xy), overwritten.  fx,fy: NDC coordinates in this winding's vertex order.
DH_HD void backward_face(const float* fx, const float* fy, int fn, float eps, const BwdMaps& m, float* grad) {
    for (int k = 0; k < 6; k++) grad[k] = 0.0f;
    if (!finite3(fx[0], fx[1], fx[2]) || !finite3(fy[0], fy[1], fy[2])) return;
    if (face_backside(fx[0], fy[0], fx[1], fy[1], fx[2], fy[2])) return;
    const int is = m.is;
    float px[3], py[3];
    for (int k = 0; k < 3; k++) {
        px[k] = ndc_to_pix(fx[k], is);
        py[k] = ndc_to_pix(fy[k], is);
    }
    for (int edge = 0; edge < 3; edge++) {
        for (int axis = 0; axis < 2; axis++) {
            Span sp;
            span_setup(px, py, edge, axis, is, sp);
            float* g_a = &grad[edge * 2 + (1 - axis)];
            float* g_b = &grad[((edge + 1) % 3) * 2 + (1 - axis)];
            for (int d0 = sp.d0_from; d0 <= sp.d0_to; d0++) {
                float d1_cross;
                int d1_in, d1_out;
                if (!span_crossing(sp, d0, is, &d1_cross, &d1_in, &d1_out)) continue;
                const int r_in = (axis == 0) ? d1_in : d0, c_in = (axis == 0) ? d0 : d1_in;
                const int r_out = (axis == 0) ? d1_out : d0, c_out = (axis == 0) ? d0 : d1_out;
                // ---- out scan: uncovered pixels beyond the edge whose gradient asks for coverage
                {
                    int d1_from, d1_to;
                    out_scan_range(sp.direction, d1_out, is, &d1_from, &d1_to);
                    const int w_from = d1_from >> 5, w_to = d1_to >> 5;
                    bool owner_known = false, owner = false;
                    for (int w = w_from; w <= w_to; w++) {
                        uint32_t bits = neg_line_word(m, axis, d0, w);
                        if (w == w_from) bits &= 0xFFFFFFFFu << (d1_from & 31);
                        if (w == w_to) bits &= 0xFFFFFFFFu >> (31 - (d1_to & 31));
                        if (!bits) continue;
                        if (!owner_known) {
                            owner = (m.fidx[r_in * is + c_in] == fn);
                            owner_known = true;
                        }
                        if (!owner) break;
                        while (bits) {
                            const int d1 = (w << 5) + ctz32(bits);
                            bits &= bits - 1;
                            const float g = (axis == 0) ? grad_at(m, d1, d0) : grad_at(m, d0, d1);
                            const float diff = (0.0f - 1.0f) * g;
                            if (diff <= 0.0f) continue;
                            float ta, tb;
                            edge_terms(diff, d0, d1, d1_cross, sp.p00, sp.p10, eps, is, &ta, &tb);
                            *g_a = *g_a - ta;
                            *g_b = *g_b - tb;
                        }
                    }
                }
                // ---- in scan: this face's own pixels, only when the pixel just outside the edge is uncovered
                if (!alpha_at(m, r_out, c_out)) {
                    int d1_from, d1_to;
                    in_scan_range(sp, d0, d1_in, is, &d1_from, &d1_to);
                    for (int d1 = d1_from; d1 <= d1_to; d1++) {
                        const int r = (axis == 0) ? d1 : d0, c = (axis == 0) ? d0 : d1;
                        if (!alpha_at(m, r, c)) continue;
                        if (!pos_at(m, r, c)) continue;
                        if (m.fidx[r * is + c] != fn) continue;
                        const float diff = (1.0f - 0.0f) * grad_at(m, r, c);
                        if (diff <= 0.0f) continue;
                        float ta, tb;
                        edge_terms(diff, d0, d1, d1_cross, sp.p00, sp.p10, eps, is, &ta, &tb);
                        *g_a = *g_a - ta;
                        *g_b = *g_b - tb;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- smoothness
// losses.py:80-84  mean((verts'[1:] - verts'[:-1])^2) and its gradient w.r.t. (T, R, s) of every frame, in
// closed form.  With verts'_b = (s v) R_b + T_b the difference of two frames is s v dR + dT, so every sum over
// the V vertices collapses onto the mesh moments m = sum v (3) and M = sum v v^T (3x3): O(1) per frame instead
// of streaming [B,V,3] three times.  double precision.
struct PairTerms { double S[3], Q[9], sse; };

// D = verts'_c - verts'_a for two poses of the same mesh, through the mesh moments m = sum v, M = sum v v^T.
DH_HD void pair_terms(const double* Ra, const double* Ta, const double* Rc, const double* Tc, double s,
                           const double* mom, double V, PairTerms& o, double* dR_out) {
    double dR[9], dT[3];
    for (int i = 0; i < 9; i++) dR[i] = Rc[i] - Ra[i];
    for (int i = 0; i < 3; i++) dT[i] = Tc[i] - Ta[i];
    const double* mv = mom;
    const double* M = mom + 3;
    double mdR[3], MdR[9];
    for (int j = 0; j < 3; j++) mdR[j] = mv[0] * dR[j] + mv[1] * dR[3 + j] + mv[2] * dR[6 + j];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            MdR[3 * i + j] = M[3 * i] * dR[j] + M[3 * i + 1] * dR[3 + j] + M[3 * i + 2] * dR[6 + j];
    double tr = 0.0;
    for (int i = 0; i < 9; i++) tr += dR[i] * MdR[i];
    o.sse = s * s * tr + 2.0 * s * (mdR[0] * dT[0] + mdR[1] * dT[1] + mdR[2] * dT[2]) +
            V * (dT[0] * dT[0] + dT[1] * dT[1] + dT[2] * dT[2]);
    for (int j = 0; j < 3; j++) o.S[j] = s * mdR[j] + V * dT[j];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) o.Q[3 * i + j] = s * s * MdR[3 * i + j] + s * mv[i] * dT[j];
    for (int i = 0; i < 9; i++) dR_out[i] = dR[i];
}

// sum_v (v R_b) . D[v]  for the pair with (dR, dT)
DH_HD double scale_term(const double* Rb, const double* dR, const double* dT, double s, const double* mom) {
    const double* mv = mom;
    const double* M = mom + 3;
    double tr = 0.0;  // tr(R_b^T M dR)
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            const double MdR = M[3 * i] * dR[j] + M[3 * i + 1] * dR[3 + j] + M[3 * i + 2] * dR[6 + j];
            tr += Rb[3 * i + j] * MdR;
        }
    double mRdT = 0.0;
    for (int j = 0; j < 3; j++) mRdT += (mv[0] * Rb[j] + mv[1] * Rb[3 + j] + mv[2] * Rb[6 + j]) * dT[j];
    return s * tr + mRdT;
}

DH_HD void load_pose(const float* rot6d, const float* trans, double* R, double* T) {
    float r6[6], Rm[9];
    for (int i = 0; i < 6; i++) r6[i] = rot6d[i];
    rot6d_to_R(r6, Rm);
    for (int i = 0; i < 9; i++) R[i] = Rm[i];
    for (int i = 0; i < 3; i++) T[i] = trans[i];
}


// Smoothness gradient (already weighted by lw_smooth) and pair loss of frame b -> st[16]:
// st[0..2] dL/dT, st[3..11] dL/dR (row-major), st[12] dL/ds, st[13] sum_v |verts'_{b+1} - verts'_b|^2.
DH_HD void smooth_terms_frame(int b, int B, const float* rot6d, const float* trans, const float* halo_prev,
                              const float* halo_next, float scale, const double* moments, int Vn, int B_total,
                              double lw_smooth, double* st) {
    for (int i = 0; i < 16; i++) st[i] = 0.0;
    if (!(lw_smooth > 0.0) || B_total < 2) return;
    const double V = (double)Vn;
    const double N = (double)(B_total - 1) * V * 3.0;
    const double coef = 2.0 * lw_smooth / N;
    const double sc = (double)scale;
    const double s = fabs(sc), sgn = (sc < 0.0) ? -1.0 : 1.0;
    double Rb[9], Tb[3], Rn[9], Tn[3], dR[9], dT[3];
    load_pose(rot6d + 6 * b, trans + 3 * b, Rb, Tb);
    PairTerms t;
    const bool has_prev = (b > 0) || (halo_prev != nullptr);
    const bool has_next = (b < B - 1) || (halo_next != nullptr);
    if (has_prev) {  // D_{b-1} = verts'_b - verts'_{b-1}
        if (b > 0) load_pose(rot6d + 6 * (b - 1), trans + 3 * (b - 1), Rn, Tn);
        else       load_pose(halo_prev, halo_prev + 6, Rn, Tn);
        pair_terms(Rn, Tn, Rb, Tb, s, moments, V, t, dR);
        for (int i = 0; i < 3; i++) dT[i] = Tb[i] - Tn[i];
        for (int i = 0; i < 3; i++) st[i] += coef * t.S[i];
        for (int i = 0; i < 9; i++) st[3 + i] += coef * t.Q[i];
        st[12] += coef * sgn * scale_term(Rb, dR, dT, s, moments);
    }
    if (has_next) {  // D_b = verts'_{b+1} - verts'_b
        if (b < B - 1) load_pose(rot6d + 6 * (b + 1), trans + 3 * (b + 1), Rn, Tn);
        else           load_pose(halo_next, halo_next + 6, Rn, Tn);
        pair_terms(Rb, Tb, Rn, Tn, s, moments, V, t, dR);
        for (int i = 0; i < 3; i++) dT[i] = Tn[i] - Tb[i];
        for (int i = 0; i < 3; i++) st[i] -= coef * t.S[i];
        for (int i = 0; i < 9; i++) st[3 + i] -= coef * t.Q[i];
        st[12] -= coef * sgn * scale_term(Rb, dR, dT, s, moments);
        st[13] = t.sse;
    }
}

// ---------------------------------------------------------------------------------------------- correspondences
// [BUILDER-DEFINED -- the reference has no such term, SURVEY.md section 0.3 / 8a]  Reprojection residual of one
// dense correspondence under the frame's pose.  Record = (X[3] canonical mesh-space point, tu, tv target position
// in ROI unit-image coordinates, w weight).  With c = (|s| X) R + T and p = K (c.x/zc, c.y/zc, 1), zc = c.z + 1e-9
// (the renderer's projection before its flip, utils/camera.py:39-57), the residual in ROI pixels is
// e = S (p - t) and the record contributes w * huber_delta(|e|) to the frame's loss.
// acc[0..2]  += dL/dT,  acc[3..11] += X_i * dL/dc_j  (so dL/dR = |s| * acc[3..11], dL/d|s| = <R, acc[3..11]>),
// acc[12] += w * huber.   Plain fp32, FMA contraction allowed (tolerance-based parity, not bit-exact).
DH_HD void corr_record(const float* rec, const float* R, const float* T, float s_abs, const float* K, float S,
                       float delta, float* acc) {
    const float X0 = rec[0], X1 = rec[1], X2 = rec[2], w = rec[5];
    const float s0 = s_abs * X0, s1 = s_abs * X1, s2 = s_abs * X2;
    const float cx = s0 * R[0] + s1 * R[3] + s2 * R[6] + T[0];
    const float cy = s0 * R[1] + s1 * R[4] + s2 * R[7] + T[1];
    const float cz = s0 * R[2] + s1 * R[5] + s2 * R[8] + T[2];
    const float iz = 1.0f / (cz + 1e-9f);
    const float x_ = cx * iz, y_ = cy * iz;
    const float eu = S * (K[0] * x_ + K[1] * y_ + K[2] - rec[3]);
    const float ev = S * (K[3] * x_ + K[4] * y_ + K[5] - rec[4]);
    const float r2 = eu * eu + ev * ev;
    float rho, f;
    if (r2 <= delta * delta) {
        rho = 0.5f * r2;
        f = 1.0f;
    } else {
        const float r = sqrtf(r2);
        rho = delta * (r - 0.5f * delta);
        f = delta / r;
    }
    const float gu = (w * f) * eu * S, gv = (w * f) * ev * S;
    const float gx_ = gu * K[0] + gv * K[3], gy_ = gu * K[1] + gv * K[4];
    const float g0 = gx_ * iz, g1 = gy_ * iz, g2 = -(gx_ * x_ + gy_ * y_) * iz;
    acc[0] += g0; acc[1] += g1; acc[2] += g2;
    acc[3] += X0 * g0; acc[4] += X0 * g1; acc[5] += X0 * g2;
    acc[6] += X1 * g0; acc[7] += X1 * g1; acc[8] += X1 * g2;
    acc[9] += X2 * g0; acc[10] += X2 * g1; acc[11] += X2 * g2;
    acc[12] += w * rho;
}

// ---------------------------------------------------------------------------------------------- exact sums
// 128-bit fixed point, value = hi + lo * 2^-64 (hi: signed integer part, floor; lo: fraction).  Adding such numbers
// is exact and associative, so a sum of doubles accumulated this way does not depend on the order or grouping of
// the terms: the gradient of the shared object scale (jointopt.py:42-46) comes out bit-identical whether the frames
// sit on one GPU or are sharded over several.  Doubles of magnitude < 2^62 convert exactly down to 2^-64.
struct Fx128 { long long hi; unsigned long long lo; };
DH_HD Fx128 fx_from_double(double x) {
    Fx128 r;
    r.hi = 0; r.lo = 0ull;
    if (!(fabs(x) < 4.0e18)) return r;   // NaN / inf / absurd magnitudes contribute nothing
    const double fl = floor(x);
    const double fr = x - fl;            // [0,1]; rounds up to 1 only for a negative x below 2^-54 in magnitude
    r.hi = (long long)fl + ((fr >= 1.0) ? 1 : 0);
    r.lo = (fr >= 1.0) ? 0ull : (unsigned long long)(fr * 18446744073709551616.0);   // exact scaling by 2^64
    return r;
}
DH_HD Fx128 fx_add(Fx128 a, Fx128 b) {
    Fx128 r;
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + ((r.lo < a.lo) ? 1 : 0);
    return r;
}
DH_HD double fx_to_double(Fx128 a) { return (double)a.hi + (double)a.lo * 5.421010862427522e-20; }  // 2^-64

// ---------------------------------------------------------------------------------------------- Adam
// torch.optim.Adam, single-tensor path, defaults betas (0.9, 0.999), eps 1e-8, no weight decay / amsgrad
// (jointopt.py:135-141).  bc1 = 1 - beta1^t, bc2s = sqrt(1 - beta2^t): python doubles, as torch computes them.
// step_size = lr / bc1 (double, then rounded to float like torch's scalar arguments).
DH_HD void adam_update(float* p, float* m, float* v, float g, float step_size, float bc2s) {
    const float b2 = (float)0.999, eps = (float)1e-8;
    const float w1 = (float)(1.0 - 0.9), w2 = (float)(1.0 - 0.999);
    const float mm = *m + w1 * (g - *m);                  // exp_avg.lerp_(grad, 1 - beta1)
    const float vv = *v * b2 + (w2 * g) * g;              // exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    const float denom = sqrtf(vv) / bc2s + eps;           // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    *p = *p + ((-step_size) * mm) / denom;                // param.addcdiv_(exp_avg, denom, value=-step_size)
    *m = mm;
    *v = vv;
}
DH_HD void adam_bias(int t, double lr, float* step_size, float* bc2s) {
    const double bc1 = 1.0 - pow(0.9, (double)t);
    const double bc2 = 1.0 - pow(0.999, (double)t);
    *step_size = (float)(lr / bc1);
    *bc2s = (float)sqrt(bc2);
}

}  // namespace dh
