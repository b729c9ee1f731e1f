// tests/emu/dh_emu.cpp -- TEST INFRASTRUCTURE ONLY.
// Host build of dynhor_b200/csrc/dh_core.h (the per-face / per-pixel arithmetic the CUDA kernels call), driven by
// serial loops that mirror the kernels' glue (strip binning, shared-memory z-buffer, bitmap staging).  The CPU
// test-suite (-m "not gpu") compares it with the oracle so that logic errors are caught without a GPU.
// The product never loads this library; the GPU tests exercise the real kernels through the C ABI.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../dynhor_b200/csrc/dh_core.h"
#include "../../dynhor_b200/csrc/dh_roi_core.h"

using namespace dh;

static const int kSH = 16;

static void load_face(const float* P, const int32_t* faces, int fn, int F, FaceSetup& fs, int* ids) {
    const int w = fn >= F;
    const int f = w ? fn - F : fn;
    const int i0 = faces[3 * f + 0], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    ids[0] = w ? i2 : i0; ids[1] = i1; ids[2] = w ? i0 : i2;
    for (int k = 0; k < 3; k++) {
        fs.x[k] = P[4 * ids[k] + 0];
        fs.y[k] = P[4 * ids[k] + 1];
        fs.z[k] = P[4 * ids[k] + 2];
    }
}

extern "C" {

void emu_rot6d_to_R(const float* r6, float* R, int B) {
    for (int b = 0; b < B; b++) rot6d_to_R(r6 + 6 * b, R + 9 * b);
}

void emu_rot6d_backward(const float* r6, const double* G, double* g6, int B) {
    for (int b = 0; b < B; b++) rot6d_backward(r6 + 6 * b, G + 9 * b, g6 + 6 * b);
}

// proj [B,V,4] from the canonical mesh and poses (k_project<true>)
void emu_project_pose(const float* verts_og, int V, const float* Rmat, const float* trans, float s_abs,
                      const float* K, float orig, int B, float* proj, float* verts_cam) {
    for (int b = 0; b < B; b++)
        for (int v = 0; v < V; v++) {
            float c[3], u, w;
            transform_vertex(verts_og + 3 * v, s_abs, Rmat + 9 * b, trans + 3 * b, c);
            project_vertex(c, K + 9 * b, orig, &u, &w);
            float* o = proj + ((size_t)b * V + v) * 4;
            o[0] = u; o[1] = w; o[2] = c[2]; o[3] = 0.f;
            if (verts_cam) memcpy(verts_cam + ((size_t)b * V + v) * 3, c, 12);
        }
}

// proj [B,V,4] from camera-space vertices (k_project<false>)
void emu_project_cam(const float* verts_cam, int V, const float* K, float orig, int B, float* proj) {
    for (int b = 0; b < B; b++)
        for (int v = 0; v < V; v++) {
            const float* c = verts_cam + ((size_t)b * V + v) * 3;
            float u, w;
            project_vertex(c, K + 9 * b, orig, &u, &w);
            float* o = proj + ((size_t)b * V + v) * 4;
            o[0] = u; o[1] = w; o[2] = c[2]; o[3] = 0.f;
        }
}

// k_setup_bin + k_raster (epilogue 1): fidx [B,is,is], alpha_bits [B,is,is/32]
void emu_raster(const float* proj, const int32_t* faces, int B, int V, int F, int is, float near, float far,
                int32_t* fidx, uint32_t* alpha_bits) {
    const int nstrips = is / kSH, wpr = is / 32;
    for (int b = 0; b < B; b++) {
        const float* P = proj + (size_t)b * V * 4;
        std::vector<std::vector<int>> bins(nstrips);
        for (int f = 0; f < F; f++)
            for (int w = 0; w < 2; w++) {
                FaceSetup fs;
                int ids[3];
                load_face(P, faces, f + w * F, F, fs, ids);
                int xl, xh, yl, yh;
                if (!face_bbox(fs.x, fs.y, is, &xl, &xh, &yl, &yh)) continue;
                for (int s = yl / kSH; s <= yh / kSH; s++) bins[s].push_back(f + w * F);
            }
        for (int strip = 0; strip < nstrips; strip++) {
            const int row0 = strip * kSH;
            std::vector<unsigned long long> zbuf((size_t)kSH * is, DH_ZKEY_EMPTY);
            // reversed order on purpose: the result must not depend on the order of bin entries
            for (int e = (int)bins[strip].size() - 1; e >= 0; e--) {
                const int fn = bins[strip][e];
                FaceSetup fs;
                int ids[3];
                load_face(P, faces, fn, F, fs, ids);
                if (!face_bbox(fs.x, fs.y, is, &fs.x_lo, &fs.x_hi, &fs.y_lo, &fs.y_hi)) continue;
                face_inverse(fs, is);
                const int r_lo = fs.y_lo > row0 ? fs.y_lo : row0;
                const int r_hi = fs.y_hi < row0 + kSH - 1 ? fs.y_hi : row0 + kSH - 1;
                for (int yi = r_lo; yi <= r_hi; yi++) {
                    const float yp = pix_to_ndc(yi, is);
                    int xa, xb;
                    row_span(fs, yp, is, fs.x_lo, fs.x_hi, &xa, &xb);
                    for (int xi = xa; xi <= xb; xi++) {
                        const float xp = pix_to_ndc(xi, is);
                        if (!pixel_inside(fs, xp, yp)) continue;
                        float zp;
                        if (!pixel_depth(fs, xi, yi, near, far, &zp)) continue;
                        const unsigned long long key = zkey(zp, fn);
                        unsigned long long& cell = zbuf[(size_t)(yi - row0) * is + xi];
                        if (key < cell) cell = key;
                    }
                }
            }
            for (int i = 0; i < kSH * is; i++) {
                const unsigned long long key = zbuf[i];
                const bool cov = key != DH_ZKEY_EMPTY;
                const int r = row0 + i / is, c = i % is;
                fidx[((size_t)b * is + r) * is + c] = cov ? (int32_t)(uint32_t)(key & 0xFFFFFFFFull) : -1;
                uint32_t& word = alpha_bits[((size_t)b * is + r) * wpr + (c >> 5)];
                if ((c & 31) == 0) word = 0;
                if (cov) word |= 1u << (c & 31);
            }
        }
    }
}

// Hit statistics of the rasteriser with deferred depth, faces taken in the kernel's pass order (given windings,
// then reversed), sequentially (tools/raster_stats.py).  out[8]: bin entries pass 0 / pass 1, hits, pre-test
// rejects, deferred installs, exact evaluations, deferred resolutions, not deferrable faces.
void emu_raster_stats(const float* proj, const int32_t* faces, int V, int F, int is, float near, float far,
                      int order, long long* out) {
    std::vector<float> depth((size_t)is * is, 0.f);
    std::vector<int> state((size_t)is * is, 0);  // 0 empty, 1 deferred, 2 exact
    long long ent[2] = {0, 0}, hits = 0, rej = 0, defer = 0, exact = 0, resolve = 0, nodefer = 0;
    for (int pi = 0; pi < 2; pi++) {
        const int pass = pi ^ order;
        for (int f = 0; f < F; f++) {
            const int fn = f + pass * F;
            FaceSetup fs;
            int ids[3];
            load_face(proj, faces, fn, F, fs, ids);
            if (!face_bbox(fs.x, fs.y, is, &fs.x_lo, &fs.x_hi, &fs.y_lo, &fs.y_hi)) continue;
            ent[pass]++;
            face_inverse(fs, is);
            const float zmin = fminf(fs.z[0], fminf(fs.z[1], fs.z[2])), zmax = fmaxf(fs.z[0], fmaxf(fs.z[1], fs.z[2]));
            const float zlo = zmin > 0.f ? zmin * (1.0f - 1e-5f) : -3.0e38f, zhi = zmax * (1.0f + 1e-5f);
            bool df = zmin > 0.f && near < zlo && zhi < far;
            for (int k = 0; k < 3; k++)
                df = df && fabsf(fs.inv[3 * k]) <= 1e3f && fabsf(fs.inv[3 * k + 1]) <= 1e3f && fabsf(fs.inv[3 * k + 2]) <= 1e6f;
            if (!df) nodefer++;
            for (int yi = fs.y_lo; yi <= fs.y_hi; yi++) {
                const float yp = pix_to_ndc(yi, is);
                for (int xi = fs.x_lo; xi <= fs.x_hi; xi++) {
                    if (!pixel_inside(fs, pix_to_ndc(xi, is), yp)) continue;
                    hits++;
                    const size_t c = (size_t)yi * is + xi;
                    if (state[c] != 0 && depth[c] < zlo) { rej++; continue; }
                    if (state[c] == 0 && df) { state[c] = 1; depth[c] = zhi; defer++; continue; }
                    float zp;
                    exact++;
                    if (!pixel_depth(fs, xi, yi, near, far, &zp)) continue;
                    if (state[c] == 1) { resolve++; state[c] = 2; depth[c] = fminf(depth[c], zp); }
                    else if (state[c] == 0 || zp < depth[c]) { state[c] = 2; depth[c] = zp; }
                }
            }
        }
    }
    long long o[8] = {ent[0], ent[1], hits, rej, defer, exact, resolve, nodefer};
    memcpy(out, o, sizeof(o));
}

// k_raster<true> epilogue 2 for whole frames: integer loss sums, dL/drend and its sign bitmaps
void emu_loss_epilogue(const uint32_t* alpha_bits, const int8_t* mask_tri, int B, int S, int aa, float gcoef,
                       int32_t* loss_counts, float* gpool, uint32_t* pos_pool, uint32_t* neg_pool, float* rend) {
    const int is = aa ? 2 * S : S, wpr = is / 32, wprp = (S + 31) / 32;
    for (int b = 0; b < B; b++) {
        int sse = 0, inter = 0, uni = 0;
        for (int yo = 0; yo < S; yo++)
            for (int x = 0; x < S; x++) {
                int pop;
                if (aa) {
                    const int r1 = is - 1 - 2 * yo, r0 = r1 - 1;
                    const uint32_t w0 = alpha_bits[((size_t)b * is + r0) * wpr + (x >> 4)];
                    const uint32_t w1 = alpha_bits[((size_t)b * is + r1) * wpr + (x >> 4)];
                    const int sh = (2 * x) & 31;
                    pop = __builtin_popcount((w0 >> sh) & 3u) + __builtin_popcount((w1 >> sh) & 3u);
                } else {
                    const int r = is - 1 - yo;
                    pop = 4 * (int)((alpha_bits[((size_t)b * is + r) * wpr + (x >> 5)] >> (x & 31)) & 1u);
                }
                const size_t o = ((size_t)b * S + yo) * S + x;
                if (rend) rend[o] = (float)pop * 0.25f;
                if (!mask_tri) continue;
                const int m = mask_tri[o];
                const int keep = m >= 0, ref = m > 0;
                const int k = keep ? pop - 4 * ref : 0;
                sse += k * k;
                inter += ref ? pop : 0;
                uni += 4 * ref + (keep ? pop : 0) - (ref ? pop : 0);
                gpool[o] = gcoef * ((float)k * 0.5f);
                uint32_t& pw = pos_pool[((size_t)b * S + yo) * wprp + (x >> 5)];
                uint32_t& nw = neg_pool[((size_t)b * S + yo) * wprp + (x >> 5)];
                if ((x & 31) == 0) { pw = 0; nw = 0; }
                if (k > 0) pw |= 1u << (x & 31);
                if (k < 0) nw |= 1u << (x & 31);
            }
        if (loss_counts) {
            loss_counts[4 * b + 0] = sse;
            loss_counts[4 * b + 1] = inter;
            loss_counts[4 * b + 2] = uni;
            loss_counts[4 * b + 3] = 0;
        }
    }
}

// k_grad_signs
void emu_grad_signs(const float* g, long long ncell, uint32_t* pos_pool, uint32_t* neg_pool) {
    for (long long i = 0; i < ncell; i++) {
        if ((i & 31) == 0) { pos_pool[i >> 5] = 0; neg_pool[i >> 5] = 0; }
        if (g[i] > 0.f) pos_pool[i >> 5] |= 1u << (i & 31);
        if (g[i] < 0.f) neg_pool[i >> 5] |= 1u << (i & 31);
    }
}

// k_backward: grad_faces [B,2F,3,2] (NDC x,y gradient of every face vertex) and, when verts_cam != NULL,
// the scattered camera-space vertex gradient grad_verts [B,V,3].
void emu_backward(const float* proj, const int32_t* faces, const int32_t* fidx, const uint32_t* alpha_bits,
                  const float* gpool, const uint32_t* pos_pool, const uint32_t* neg_pool, int B, int V, int F, int S,
                  int aa, float eps, float* grad_faces, const float* verts_cam, const float* K, float orig,
                  float* grad_verts) {
    const int is = aa ? 2 * S : S, wpr = is / 32, wprp = (S + 31) / 32;
    if (grad_verts) memset(grad_verts, 0, sizeof(float) * (size_t)B * V * 3);
    for (int b = 0; b < B; b++) {
        std::vector<uint32_t> s_alpha((size_t)is * wpr), s_neg((size_t)is * wpr), s_negT((size_t)is * wpr, 0u);
        const uint32_t* ga = alpha_bits + (size_t)b * is * wpr;
        const uint32_t* gn = neg_pool + (size_t)b * S * wprp;
        for (int i = 0; i < is * wpr; i++) {
            const uint32_t a = ga[i];
            const int r = i / wpr, w = i - r * wpr;
            const int rf = is - 1 - r;
            uint32_t nb;
            if (aa) {
                const uint32_t pw = gn[(rf >> 1) * wprp + (w >> 1)];
                nb = spread16((w & 1) ? (pw >> 16) : pw);
            } else {
                nb = gn[rf * wprp + w];
            }
            s_alpha[i] = a;
            s_neg[i] = ~a & nb;
        }
        for (int r = 0; r < is; r++)
            for (int c = 0; c < is; c++)
                if ((s_neg[(size_t)r * wpr + (c >> 5)] >> (c & 31)) & 1u) s_negT[(size_t)c * wpr + (r >> 5)] |= 1u << (r & 31);
        BwdMaps m;
        m.alpha = s_alpha.data(); m.neg = (b & 1) ? s_neg.data() : nullptr; m.negT = s_negT.data();
        m.neg_pool = gn;  // odd frames use the stored row-major bitmap, even frames the on-the-fly derivation
        m.row_lo = m.row_hi = m.col_lo = m.col_hi = nullptr;
        m.pos_pool = pos_pool + (size_t)b * S * wprp;
        m.gpool = gpool + (size_t)b * S * S;
        m.fidx = fidx + (size_t)b * is * is;
        m.is = is; m.S = S; m.aa = aa; m.wpr = wpr; m.wpr_pool = wprp;
        m.gscale = aa ? 0.25f : 1.0f;
        const float* P = proj + (size_t)b * V * 4;
        for (int fn = 0; fn < 2 * F; fn++) {
            FaceSetup fs;
            int ids[3];
            load_face(P, faces, fn, F, fs, ids);
            float g[6];
            backward_face(fs.x, fs.y, fn, eps, m, g);
            memcpy(grad_faces + ((size_t)b * 2 * F + fn) * 6, g, sizeof(g));
            if (!grad_verts) continue;
            for (int k = 0; k < 3; k++) {
                if (g[2 * k] == 0.f && g[2 * k + 1] == 0.f) continue;
                float gc[3];
                project_vertex_backward(verts_cam + ((size_t)b * V + ids[k]) * 3, K + 9 * b, orig, g[2 * k],
                                        g[2 * k + 1], gc);
                for (int j = 0; j < 3; j++) grad_verts[((size_t)b * V + ids[k]) * 3 + j] += gc[j];
            }
        }
    }
}

// Work counters of the edge-scan backward of one frame (tools/bwd_stats.py): how many items, crossings, tasks,
// bitmap words and contributing pixels the kernel's stages see.  out[16] (long long).
void emu_backward_stats(const float* proj, const int32_t* faces, const int32_t* fidx, const uint32_t* alpha_bits,
                        const uint32_t* neg_pool, int V, int F, int S, int aa, long long* out, int* span_len) {
    const int is = aa ? 2 * S : S, wpr = is / 32, wprp = (S + 31) / 32;
    std::vector<uint32_t> s_neg((size_t)is * wpr), s_negT((size_t)is * wpr, 0u);
    std::vector<int> rlo(is, is), rhi(is, -1), clo(is, is), chi(is, -1);
    long long n_neg = 0;
    for (int i = 0; i < is * wpr; i++) {
        const int r = i / wpr, w = i - r * wpr;
        s_neg[i] = neg_row_word(alpha_bits, neg_pool, is, aa, wpr, wprp, r, w);
    }
    for (int r = 0; r < is; r++)
        for (int c = 0; c < is; c++)
            if ((s_neg[(size_t)r * wpr + (c >> 5)] >> (c & 31)) & 1u) {
                s_negT[(size_t)c * wpr + (r >> 5)] |= 1u << (r & 31);
                n_neg++;
                if (c < rlo[r]) rlo[r] = c;
                if (c > rhi[r]) rhi[r] = c;
                if (r < clo[c]) clo[c] = r;
                if (r > chi[c]) chi[c] = r;
            }
    std::vector<char> owned((size_t)2 * F, 0);
    for (int i = 0; i < is * is; i++) if (fidx[i] >= 0) owned[fidx[i]] = 1;
    long long items = 0, crossings = 0, t_out = 0, t_out_owner = 0, words = 0, words_nz = 0, pairs = 0, t_in = 0,
              in_px = 0, in_pairs = 0, max_pairs_task = 0, span_iters = 0, front = 0;
    for (int fn = 0; fn < 2 * F; fn++) {
        FaceSetup fs;
        int ids[3];
        load_face(proj, faces, fn, F, fs, ids);
        if (!finite3(fs.x[0], fs.x[1], fs.x[2]) || !finite3(fs.y[0], fs.y[1], fs.y[2])) continue;
        if (face_backside(fs.x[0], fs.y[0], fs.x[1], fs.y[1], fs.x[2], fs.y[2])) continue;
        front++;
        if (!owned[fn]) continue;
        items++;
        float px[3], py[3];
        for (int k = 0; k < 3; k++) { px[k] = ndc_to_pix(fs.x[k], is); py[k] = ndc_to_pix(fs.y[k], is); }
        for (int edge = 0; edge < 3; edge++)
            for (int axis = 0; axis < 2; axis++) {
                Span sp;
                span_setup(px, py, edge, axis, is, sp);
                span_iters++;
                if (span_len) span_len[(items - 1) * 6 + edge * 2 + axis] = std::max(0, sp.d0_to - sp.d0_from + 1);
                for (int d0 = sp.d0_from; d0 <= sp.d0_to; d0++) {
                    span_iters++;
                    float d1_cross;
                    int d1_in, d1_out;
                    if (!span_crossing(sp, d0, is, &d1_cross, &d1_in, &d1_out)) continue;
                    crossings++;
                    int from, to;
                    out_scan_range(sp.direction, d1_out, is, &from, &to);
                    const int lo = axis == 0 ? clo[d0] : rlo[d0], hi = axis == 0 ? chi[d0] : rhi[d0];
                    const int r_in = (axis == 0) ? d1_in : d0, c_in = (axis == 0) ? d0 : d1_in;
                    const int r_out = (axis == 0) ? d1_out : d0, c_out = (axis == 0) ? d0 : d1_out;
                    if (std::max(from, lo) <= std::min(to, hi)) {
                        t_out++;
                        if (fidx[r_in * is + c_in] == fn) {
                            t_out_owner++;
                            from = std::max(from, lo); to = std::min(to, hi);
                            long long np = 0;
                            for (int w = from >> 5; w <= (to >> 5); w++) {
                                uint32_t bits = axis == 0 ? s_negT[(size_t)d0 * wpr + w] : s_neg[(size_t)d0 * wpr + w];
                                if (w == (from >> 5)) bits &= 0xFFFFFFFFu << (from & 31);
                                if (w == (to >> 5)) bits &= 0xFFFFFFFFu >> (31 - (to & 31));
                                words++;
                                if (bits) words_nz++;
                                np += __builtin_popcount(bits);
                            }
                            pairs += np;
                            if (np > max_pairs_task) max_pairs_task = np;
                        }
                    }
                    if (!((alpha_bits[r_out * wpr + (c_out >> 5)] >> (c_out & 31)) & 1u)) {
                        t_in++;
                        int f2, t2;
                        in_scan_range(sp, d0, d1_in, is, &f2, &t2);
                        in_px += std::max(0, t2 - f2 + 1);
                    }
                }
            }
    }
    long long o[16] = {front, items, span_iters, crossings, t_out, t_out_owner, words, words_nz, pairs, t_in, in_px,
                       in_pairs, max_pairs_task, n_neg, 0, 0};
    memcpy(out, o, sizeof(o));
}

// dh_roi.cu on the host: tight bounds, boxes, ROIAlign crops of the object / occluder bit masks and of the image
// status[b] = 1 if the object mask of frame b is empty (the reference's np.min raises there)
void emu_roi_process(const uint8_t* obj_bits, const uint8_t* hand_bits, const uint8_t* images_hwc, int B, int H, int W,
                     int S, float pad, float expansion, float* bbox, float* square_bbox, uint8_t* crop_mask,
                     float* target, float* crop_image, int32_t* status) {
    for (int b = 0; b < B; b++) {
        const uint8_t* ob = obj_bits + (size_t)b * H * W;
        const uint8_t* hb = hand_bits ? hand_bits + (size_t)b * H * W : nullptr;
        int r0 = H, r1 = -1, c0 = W, c1 = -1;
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
                if (ob[(size_t)y * W + x]) {
                    r0 = std::min(r0, y); r1 = std::max(r1, y); c0 = std::min(c0, x); c1 = std::max(c1, x);
                }
        status[b] = r1 < 0;
        if (r1 < 0) continue;
        float xyxy[4];
        dh::roi_boxes(r0, r1, c0, c1, H, W, pad, expansion, bbox + 4 * b, square_bbox + 4 * b, xyxy);
        const dh::RoiGeom g = dh::roi_geom(xyxy, S);
        for (int ph = 0; ph < S; ph++)
            for (int pw = 0; pw < S; pw++) {
                const float vo = dh::roi_align_cell(g, ph, pw, H, W, [&](int y, int x) { return (float)ob[(size_t)y * W + x]; });
                const bool obit = vo >= 0.5f;
                bool hbit = false;
                if (hb) hbit = dh::roi_align_cell(g, ph, pw, H, W, [&](int y, int x) { return (float)hb[(size_t)y * W + x]; }) >= 0.5f;
                const size_t o = ((size_t)b * S + ph) * S + pw;
                crop_mask[o] = obit;
                target[o] = dh::target_value(obit, hbit);
                if (images_hwc && crop_image) {
                    const uint8_t* im = images_hwc + (size_t)b * H * W * 3;
                    float v[3] = {1.0f, 1.0f, 1.0f};
                    if (obit)
                        dh::roi_align_cell3(g, ph, pw, H, W, [&](int y, int x, float* px) {
                            for (int c = 0; c < 3; c++) px[c] = (float)((double)im[((size_t)y * W + x) * 3 + c] / 255.0);
                        }, v);
                    for (int c = 0; c < 3; c++) crop_image[(((size_t)b * 3 + c) * S + ph) * S + pw] = v[c];
                }
            }
    }
}

// k_pose_prep: st [B,16]
void emu_smooth_terms(const float* rot6d, const float* trans, const float* halo_prev, const float* halo_next,
                      float scale, const double* moments, int V, int B, int B_total, double lw_smooth, double* st) {
    for (int b = 0; b < B; b++)
        smooth_terms_frame(b, B, rot6d, trans, halo_prev, halo_next, scale, moments, V, B_total, lw_smooth,
                           st + (size_t)16 * b);
}

void emu_adam(float* p, const float* g, float* m, float* v, long long n, double lr, int t) {
    float step_size, bc2s;
    adam_bias(t, lr, &step_size, &bc2s);
    for (long long i = 0; i < n; i++) adam_update(&p[i], &m[i], &v[i], g[i], step_size, bc2s);
}

// k_corr (dh_corr.cu): per-frame sums of the correspondence term, thread-strided like the kernel
// (256 "threads", each walking its records in order, then summed in thread order).  sums [B,16].
void emu_corr_frames(const float* records, int B, int C, const float* Rmat, const float* trans, float s_abs,
                     const float* K, float S, float delta, float* sums) {
    for (int b = 0; b < B; b++) {
        std::vector<float> acc(256 * 13, 0.f);
        for (int c = 0; c < C; c++)
            corr_record(records + ((size_t)b * C + c) * 6, Rmat + 9 * b, trans + 3 * b, s_abs, K + 9 * b, S, delta,
                        &acc[(c % 256) * 13]);
        for (int j = 0; j < 16; j++) sums[b * 16 + j] = 0.f;
        for (int t = 0; t < 256; t++)
            for (int j = 0; j < 13; j++) sums[b * 16 + j] += acc[t * 13 + j];
    }
}

// k_finalize's exact scale-gradient sum: out[0] = hi, out[1] = lo of the sum of x[0..n), *val = its double value
void emu_fx_sum(const double* x, int n, unsigned long long* out2, double* val) {
    Fx128 t;
    t.hi = 0; t.lo = 0ull;
    for (int i = 0; i < n; i++) t = fx_add(t, fx_from_double(x[i]));
    out2[0] = (unsigned long long)t.hi;
    out2[1] = t.lo;
    *val = fx_to_double(t);
}


// The backward's "one crossing per lane" enumeration of a batch (k_backward<lists>, dh_jointopt.cu), emulated for a
// warp of 32 lanes with the SAME helpers (span_setup, span_info, span_mark, span_of_lane, span_info_d0): the span list
// is built in (lane, span) order from prefix sums, then the crossings 0 ... T-1 are handed out 32 per step with the
// mark / popcount trick.  px, py: [32][3] pixel coordinates of the faces (first n_have lanes valid).
// flat[3 i ..] = (lane slot, edge * 2 + axis, scan line) of crossing i as the kernel's lanes see it;
// ref[3 i ..]  = the same from the plain nested loops (lane, span, scan line).  Returns T (<= cap), or -1.
int emu_span_list(const float* px, const float* py, int n_have, int is, int32_t* flat, int32_t* ref, int cap) {
    std::vector<uint16_t> start16;
    std::vector<uint32_t> info;
    uint32_t tb = 0;
    int nref = 0;
    for (int lane = 0; lane < 32 && lane < n_have; lane++) {
        for (int k = 0; k < 6; k++) {
            Span t;
            span_setup(px + 3 * lane, py + 3 * lane, k >> 1, k & 1, is, t);
            const int l = t.d0_to - t.d0_from + 1;
            if (l <= 0) continue;
            start16.push_back((uint16_t)tb);
            info.push_back(span_info(lane, k, t.d0_from, 0 < t.direction, tb));
            for (int d0 = t.d0_from; d0 <= t.d0_to; d0++) {
                if (nref >= cap) return -1;
                ref[3 * nref + 0] = lane; ref[3 * nref + 1] = k; ref[3 * nref + 2] = d0;
                nref++;
            }
            tb += (uint32_t)l;
        }
    }
    const int T = (int)tb, NS = (int)info.size();
    if (T != nref || NS > 192) return -1;
    int s0 = 0;
    for (int base = 0; base < T; base += 32) {
        uint32_t M = 0;
        for (int lane = 0; lane < 32; lane++) {
            const int j = s0 + 1 + lane;
            if (j < NS) M |= span_mark(start16[j], base);
        }
        for (int lane = 0; lane < 32; lane++) {
            const int idx = base + lane;
            if (idx >= T) break;
            const int span = span_of_lane(s0, M, lane);
            if (span < 0 || span >= NS) return -1;
            const uint32_t inf = info[span];
            flat[3 * idx + 0] = (int)(inf & 31u);
            flat[3 * idx + 1] = (int)((inf >> 5) & 7u);
            flat[3 * idx + 2] = span_info_d0(inf, idx);
        }
        s0 += popc32(M);
    }
    return T;
}

}  // extern "C"
