/*
 * oracle/nmr_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never a product path).
 *
 * CPU restatement of the silhouette rasteriser that the reference calls at
 *   ObjTracker/utils/losses.py:36-40,68     (nr.renderer.Renderer(...)(verts, faces, mode="silhouettes"))
 *   ObjTracker/pose_initializtion.py:98-105,146-147,160
 * The arithmetic lives in the third-party CUDA extension `neural_renderer`
 * (ObjTracker/requirements.txt:8, git+https://github.com/hassony2/multiperson.git@master#subdirectory=neural_renderer,
 * UNPINNED branch ref, NOT vendored under /root/reference, not installed, no network).
 * This file restates that package's published algorithm (Kato et al., "Neural 3D Mesh Renderer",
 * PyTorch port: rasterize_cuda_kernel.cu -- forward_face_index_map kernels 1/2, forward_alpha_map,
 * backward_pixel_map) as summarised in SURVEY.md Appendix A.1/A.2.
 *
 * PARITY UNPINNED at this boundary: the reference ships no tests, golden vectors or fixtures
 * (SURVEY.md section 4 / 8c), so this oracle is the normative definition of the rasteriser for this build.
 * Everything around it (projection, losses, Adam) IS pinned against the reference's own Python
 * (tests/golden/make_golden.py).
 *
 * Numerics: plain IEEE fp32, same operation order as the published kernels, with the
 * double-typed literals of the published source kept where they matter. Must be compiled
 * with -ffp-contract=off and without -ffast-math (oracle/Makefile does that), so no FMA is
 * formed: the reading of the source that does not depend on a particular compiler.
 *
 * Same brute-force structure as the third-party kernels on purpose (one pixel loops over all
 * faces; one face loops over its edge scan lines): it doubles as the "reference on CPU"
 * baseline. OpenMP over pixels / faces only.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* CUDA's double->int conversion (cvt.rzi.s32.f64): truncate, saturate, NaN -> 0. */
static inline int d2i(double v) {
    if (v != v) return 0;
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return (-2147483647 - 1);
    return (int)v;
}
static inline int f2i(float v) { return d2i((double)v); }

static inline int is_backside(const float* face) {
    /* (y2-y0)(x1-x0) < (y1-y0)(x2-x0) */
    float a = (face[7] - face[1]) * (face[3] - face[0]);
    float b = (face[4] - face[1]) * (face[6] - face[0]);
    return a < b;
}

/* Kernel 1 of the forward pass: per-face inverse of [[x0,x1,x2],[y0,y1,y2],[1,1,1]] in pixel coords.
 * faces: [B, NF, 3, 3] (x,y in NDC [-1,1], z depth).  faces_inv: [B, NF, 9], zero for back faces. */
static void face_inverse(const float* face, float* face_inv_g, int is) {
    float p[3][2];
    for (int num = 0; num < 3; num++)
        for (int dim = 0; dim < 2; dim++) {
            float t = face[3 * num + dim] * (float)is;
            t = t + (float)is;
            t = t - 1.0f;
            p[num][dim] = (float)(0.5 * (double)t);
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
    for (int k = 0; k < 9; k++) face_inv_g[k] = fi[k] / den;
}

/* Forward: face_index_map [B,is,is] (-1 = none), weight_map [B,is,is,3], depth_map [B,is,is] (far where none),
 * alpha_map [B,is,is] in {0,1}.  Rows are in the rasteriser's own (un-flipped) order. */
void nmr_forward(const float* faces, int B, int NF, int is, float near, float far,
                 int32_t* face_index_map, float* weight_map, float* depth_map, float* alpha_map) {
    size_t nface = (size_t)B * NF;
    float* faces_inv = (float*)calloc(nface * 9, sizeof(float));
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)nface; i++) {
        const float* face = faces + (size_t)i * 9;
        if (is_backside(face)) continue;
        face_inverse(face, faces_inv + (size_t)i * 9, is);
    }
    long npix = (long)B * is * is;
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < npix; i++) {
        const int bn = (int)(i / ((long)is * is));
        const int pn = (int)(i % ((long)is * is));
        const int yi = pn / is;
        const int xi = pn % is;
        const float yp = (float)((2. * yi + 1 - is) / is);
        const float xp = (float)((2. * xi + 1 - is) / is);
        const float* fbase = faces + (size_t)bn * NF * 9;
        const float* ibase = faces_inv + (size_t)bn * NF * 9;
        float depth_min = far;
        int face_index_min = -1;
        float weight_min[3] = {0.f, 0.f, 0.f};
        for (int fn = 0; fn < NF; fn++) {
            const float* face = fbase + (size_t)fn * 9;
            const float* face_inv = ibase + (size_t)fn * 9;
            if (is_backside(face)) continue;
            if (((yp - face[1]) * (face[3] - face[0]) < (xp - face[0]) * (face[4] - face[1])) ||
                ((yp - face[4]) * (face[6] - face[3]) < (xp - face[3]) * (face[7] - face[4])) ||
                ((yp - face[7]) * (face[0] - face[6]) < (xp - face[6]) * (face[1] - face[7])))
                continue;
            float w[3];
            for (int k = 0; k < 3; k++) {
                float t = face_inv[3 * k + 0] * (float)xi;
                t = t + face_inv[3 * k + 1] * (float)yi;
                w[k] = t + face_inv[3 * k + 2];
            }
            float w_sum = 0.f;
            for (int k = 0; k < 3; k++) {
                w[k] = (float)fmin(fmax((double)w[k], 0.), 1.);
                w_sum = w_sum + w[k];
            }
            for (int k = 0; k < 3; k++) w[k] = w[k] / w_sum;
            float s = w[0] / face[2];
            s = s + w[1] / face[5];
            s = s + w[2] / face[8];
            const float zp = (float)(1. / (double)s);
            if (zp <= near || far <= zp) continue;
            if (zp < depth_min) {
                depth_min = zp;
                face_index_min = fn;
                for (int k = 0; k < 3; k++) weight_min[k] = w[k];
            }
        }
        if (0 <= face_index_min) {
            depth_map[i] = depth_min;
            face_index_map[i] = face_index_min;
            for (int k = 0; k < 3; k++) weight_map[3 * i + k] = weight_min[k];
            alpha_map[i] = 1.f;
        } else {
            depth_map[i] = far;
            face_index_map[i] = -1;
            for (int k = 0; k < 3; k++) weight_map[3 * i + k] = 0.f;
            alpha_map[i] = 0.f;
        }
    }
    free(faces_inv);
}

/* Backward of the alpha map w.r.t. face vertices (x,y only; z gets 0): the hand-designed
 * "edge scan" pseudo-gradient.  grad_faces: [B,NF,3,3], fully overwritten. */
void nmr_backward(const float* faces, const int32_t* face_index_map, const float* alpha_map,
                  const float* grad_alpha_map, float* grad_faces, int B, int NF, int is, float eps) {
    long nface = (long)B * NF;
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < nface; i++) {
        const int bn = (int)(i / NF);
        const int fn = (int)(i % NF);
        const float* face = faces + (size_t)i * 9;
        float grad_face[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float* out = grad_faces + (size_t)i * 9;
        if (is_backside(face)) {
            for (int k = 0; k < 9; k++) out[k] = 0.f;
            continue;
        }
        const size_t mbase = (size_t)bn * is * is;
        for (int edge_num = 0; edge_num < 3; edge_num++) {
            int pi[3];
            float pp[3][2];
            for (int num = 0; num < 3; num++) pi[num] = (edge_num + num) % 3;
            for (int num = 0; num < 3; num++)
                for (int dim = 0; dim < 2; dim++) {
                    float t = face[3 * pi[num] + dim] * (float)is;
                    t = t + (float)is;
                    t = t - 1.0f;
                    pp[num][dim] = (float)(0.5 * (double)t);
                }
            for (int axis = 0; axis < 2; axis++) {
                float p[3][2];
                for (int num = 0; num < 3; num++)
                    for (int dim = 0; dim < 2; dim++) p[num][dim] = pp[num][(dim + axis) % 2];
                int direction;
                if (axis == 0) direction = (p[0][0] < p[1][0]) ? -1 : 1;
                else           direction = (p[0][0] < p[1][0]) ? 1 : -1;
                const int d0_from = d2i(fmax(ceil((double)fminf(p[0][0], p[1][0])), 0.));
                const int d0_to = d2i(fmin((double)fmaxf(p[0][0], p[1][0]), is - 1.));
                for (int d0 = d0_from; d0 <= d0_to; d0++) {
                    float d1_cross = (p[1][1] - p[0][1]) / (p[1][0] - p[0][0]);
                    d1_cross = d1_cross * ((float)d0 - p[0][0]);
                    d1_cross = d1_cross + p[0][1];
                    int d1_in;
                    if (0 < direction) d1_in = f2i(floorf(d1_cross));
                    else               d1_in = f2i(ceilf(d1_cross));
                    const int d1_out = d1_in + direction;
                    if (d1_in < 0 || is <= d1_in) continue;
                    if (d1_out < 0 || is <= d1_out) continue;
                    size_t map_index_in, map_index_out;
                    if (axis == 0) {
                        map_index_in = mbase + (size_t)d1_in * is + d0;
                        map_index_out = mbase + (size_t)d1_out * is + d0;
                    } else {
                        map_index_in = mbase + (size_t)d0 * is + d1_in;
                        map_index_out = mbase + (size_t)d0 * is + d1_out;
                    }
                    const float alpha_in = alpha_map[map_index_in];
                    const float alpha_out = alpha_map[map_index_out];
                    const int map_offset = (axis == 0) ? is : 1;
                    /* out: pixels beyond the edge, up to the image border */
                    if (face_index_map[map_index_in] == fn) {
                        const int d1_limit = (0 < direction) ? is - 1 : 0;
                        int d1_from = d1_out < d1_limit ? d1_out : d1_limit;
                        if (d1_from < 0) d1_from = 0;
                        int d1_to = d1_out > d1_limit ? d1_out : d1_limit;
                        if (d1_to > is - 1) d1_to = is - 1;
                        size_t idx = (axis == 0) ? mbase + (size_t)d1_from * is + d0
                                                 : mbase + (size_t)d0 * is + d1_from;
                        for (int d1 = d1_from; d1 <= d1_to; d1++, idx += map_offset) {
                            float diff_grad = 0.f;
                            diff_grad = diff_grad + (alpha_map[idx] - alpha_in) * grad_alpha_map[idx];
                            if (diff_grad <= 0) continue;
                            if (p[1][0] != (float)d0) {
                                float t = (p[1][0] - p[0][0]) / (p[1][0] - (float)d0);
                                t = t * ((float)d1 - d1_cross);
                                float dist = (float)(((double)t * 2.) / is);
                                dist = (0 < dist) ? dist + eps : dist - eps;
                                grad_face[pi[0] * 3 + (1 - axis)] -= diff_grad / dist;
                            }
                            if (p[0][0] != (float)d0) {
                                float t = (p[1][0] - p[0][0]) / ((float)d0 - p[0][0]);
                                t = t * ((float)d1 - d1_cross);
                                float dist = (float)(((double)t * 2.) / is);
                                dist = (0 < dist) ? dist + eps : dist - eps;
                                grad_face[pi[1] * 3 + (1 - axis)] -= diff_grad / dist;
                            }
                        }
                    }
                    /* in: pixels of this face between the edge and the opposite edge */
                    {
                        float d0_cross2;
                        if (((float)d0 - p[0][0]) * ((float)d0 - p[2][0]) < 0) {
                            d0_cross2 = (p[2][1] - p[0][1]) / (p[2][0] - p[0][0]);
                            d0_cross2 = d0_cross2 * ((float)d0 - p[0][0]);
                            d0_cross2 = d0_cross2 + p[0][1];
                        } else {
                            d0_cross2 = (p[1][1] - p[2][1]) / (p[1][0] - p[2][0]);
                            d0_cross2 = d0_cross2 * ((float)d0 - p[2][0]);
                            d0_cross2 = d0_cross2 + p[2][1];
                        }
                        int d1_limit;
                        if (0 < direction) d1_limit = f2i(ceilf(d0_cross2));
                        else               d1_limit = f2i(floorf(d0_cross2));
                        int d1_from = d1_in < d1_limit ? d1_in : d1_limit;
                        if (d1_from < 0) d1_from = 0;
                        int d1_to = d1_in > d1_limit ? d1_in : d1_limit;
                        if (d1_to > is - 1) d1_to = is - 1;
                        size_t idx = (axis == 0) ? mbase + (size_t)d1_from * is + d0
                                                 : mbase + (size_t)d0 * is + d1_from;
                        for (int d1 = d1_from; d1 <= d1_to; d1++, idx += map_offset) {
                            if (face_index_map[idx] != fn) continue;
                            float diff_grad = 0.f;
                            diff_grad = diff_grad + (alpha_map[idx] - alpha_out) * grad_alpha_map[idx];
                            if (diff_grad <= 0) continue;
                            if (p[1][0] != (float)d0) {
                                float t = (p[1][0] - p[0][0]) / (p[1][0] - (float)d0);
                                t = t * ((float)d1 - d1_cross);
                                float dist = (float)(((double)t * 2.) / is);
                                dist = (0 < dist) ? dist + eps : dist - eps;
                                grad_face[pi[0] * 3 + (1 - axis)] -= diff_grad / dist;
                            }
                            if (p[0][0] != (float)d0) {
                                float t = (p[1][0] - p[0][0]) / ((float)d0 - p[0][0]);
                                t = t * ((float)d1 - d1_cross);
                                float dist = (float)(((double)t * 2.) / is);
                                dist = (0 < dist) ? dist + eps : dist - eps;
                                grad_face[pi[1] * 3 + (1 - axis)] -= diff_grad / dist;
                            }
                        }
                    }
                }
            }
        }
        for (int k = 0; k < 9; k++) out[k] = grad_face[k];
    }
}

int nmr_oracle_abi_version(void) { return 1; }
