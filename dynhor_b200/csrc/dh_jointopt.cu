// dh_jointopt.cu -- sm_100a kernels of the joint pose-optimisation hot path and their C-ABI entry points.
//
// One optimisation iteration (jointopt.py:144-160) is ten kernels, replayed as one CUDA graph (k_corr as a branch
// beside k_neg_maps; the first iteration of a call may run as two plain-stream halves, dh_jointopt_run_part):
//   k_pose_prep    6D rotation -> R (geometry.py:19-25); closed-form smoothness loss + gradient from mesh moments
//   k_corr         (dh_corr.cu, builder-defined correspondence term; only with correspondences and a positive weight)
//   k_project      (|s| v) R + T  (camera.py:204-206) and the renderer's projection (camera.py:39-62) -> NDC
//   k_setup_bin    fill_back + back-face cull + conservative bbox; faces binned to 16-row strips
//   k_raster       one CTA per (frame, strip): 64-bit (depth, face) z-buffer in shared memory, two winding passes
//                  with tile-z culling in between, then the fused epilogue: face-index map, coverage bitmap, 2x2
//                  pooling + flip, masked-L2 / IoU integer sums, dL/drend map and its sign bitmaps (losses.py:66-78)
//   k_neg_maps     per frame: the "wanted but uncovered" pixels as a transposed bitmap, first / last such pixel of
//                  every row and column, per-line lists, list-overflow flag
//   k_backward     <lists>: one CTA per (frame, face chunk): coverage bitmap + line ranges staged in shared memory,
//                  guided batches of items, one edge crossing per lane over the batch's span list, per-face
//                  edge-scan pseudo-gradient over the pixel lists, projection + rigid-transform backward, CTA
//                  reduction to 13 numbers;  <bitmaps>: the same on bitmap words, for frames whose lists overflowed
//                  (a small grid that normally ends after reading the flags) and for the composable dh_sil_backward
//   k_pose_update  a warp per frame: partial sums + smoothness gradient, Gram-Schmidt backward, two-group Adam
//                  (jointopt.py:135-141)
//   k_finalize     per-iteration loss / IoU sums -> history row, step counter, optional scale update
// Compiled with --fmad=false: see dh_core.h for the fp32 contract.
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "dh_common.h"
#include "dh_core.h"

namespace {

using namespace dh;

#ifndef DH_STRIP_ROWS
#define DH_STRIP_ROWS 16
#endif
constexpr int kSH = DH_STRIP_ROWS;  // strip height in raster rows (even: a strip holds whole 2x2 pooling cells)
constexpr int kThreads = 256;      // CTA size of the backward / elementwise kernels
#ifndef DH_RASTER_THREADS
#define DH_RASTER_THREADS 416
#endif
#ifndef DH_RASTER_EVEN
#define DH_RASTER_EVEN 1
#endif
#ifndef DH_RASTER_SPLIT1
#define DH_RASTER_SPLIT1 1  // the same for the second pass (reversed windings: mostly culled by tile-z)
#endif
#ifndef DH_RASTER_SPLIT
#define DH_RASTER_SPLIT 1   // batches per warp the last (partial) round of the first pass is cut into
#endif
#ifndef DH_TILE_Z
#define DH_TILE_Z 1
#endif
#ifndef DH_DEFER_DEPTH
#define DH_DEFER_DEPTH 1   // 1: deferred depth (see raster_hit): ~77k -> ~200 exact depth evaluations per frame, 12 % fewer
#endif                     // instructions, bit-exact.  Round 1: 1.09 vs 1.08 ms (off); round 2: 0.907 vs 0.920 ms (on)
constexpr int kRasterWarps = DH_RASTER_THREADS / 32;
constexpr int kRasterThreads = DH_RASTER_THREADS;  // 13 warps x 2 CTAs/SM: the 64 KB z-buffer strip caps CTAs/SM at 2
constexpr int kOwnedSmemWords = 1024;
constexpr int kMaxIS = 512;        // largest raster resolution (bitmaps + z-buffer strip must fit shared memory)

__host__ __device__ inline int raster_size(const dh_sil& s) { return s.aa ? 2 * s.S : s.S; }

// ------------------------------------------------------------------------------------------------ small ops
__global__ void k_rot6d_to_matrix(const float* __restrict__ rot6d, float* __restrict__ R, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float r6[6], Rm[9];
    for (int i = 0; i < 6; i++) r6[i] = rot6d[6 * b + i];
    rot6d_to_R(r6, Rm);
    for (int i = 0; i < 9; i++) R[9 * b + i] = Rm[i];
}

__global__ void k_transform_verts(const float* __restrict__ verts, const float* __restrict__ R,
                                  const float* __restrict__ T, const float* __restrict__ scale,
                                  float* __restrict__ out, int V) {
    const int b = blockIdx.y;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    float Rm[9], Tm[3], vv[3], o[3];
    for (int i = 0; i < 9; i++) Rm[i] = R[9 * b + i];
    for (int i = 0; i < 3; i++) Tm[i] = T[3 * b + i];
    for (int i = 0; i < 3; i++) vv[i] = verts[3 * v + i];
    transform_vertex(vv, fabsf(scale[0]), Rm, Tm, o);
    float* dst = out + ((size_t)b * V + v) * 3;
    dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
}

__global__ void k_masks_prepare(const float* __restrict__ m, int8_t* __restrict__ tri,
                                unsigned long long* __restrict__ keep_count, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned int keep = 0;
    for (; i < n; i += stride) {
        const float v = m[i];
        const int8_t t = (v > 0.0f) ? 1 : ((v >= 0.0f) ? 0 : -1);
        tri[i] = t;
        keep += (t >= 0);
    }
    for (int o = 16; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
    if ((threadIdx.x & 31) == 0 && keep) atomicAdd(keep_count, (unsigned long long)keep);
}

__global__ void k_mesh_moments(const float* __restrict__ verts, int V, double* __restrict__ out12) {
    __shared__ double sm[kThreads / 32][12];
    double a[12];
    for (int i = 0; i < 12; i++) a[i] = 0.0;
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        const double x[3] = {verts[3 * v], verts[3 * v + 1], verts[3 * v + 2]};
        for (int i = 0; i < 3; i++) {
            a[i] += x[i];
            for (int j = 0; j < 3; j++) a[3 + 3 * i + j] += x[i] * x[j];
        }
    }
    for (int i = 0; i < 12; i++)
        for (int o = 16; o > 0; o >>= 1) a[i] += __shfl_xor_sync(0xffffffffu, a[i], o);
    if ((threadIdx.x & 31) == 0)
        for (int i = 0; i < 12; i++) sm[threadIdx.x >> 5][i] = a[i];
    __syncthreads();
    if (threadIdx.x < 12) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += sm[w][threadIdx.x];
        out12[threadIdx.x] = s;
    }
}

__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, long long n, float step_size, float bc2s) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    adam_update(&p[i], &m[i], &v[i], g[i], step_size, bc2s);
}

// ------------------------------------------------------------------------------------------------ projection
// FROM_POSE: verts = canonical mesh [V,3], transformed by R[b], T[b], |s|.  else: verts = camera space [B,V,3].
// Block (0, b) also clears the frame's bin counters and loss counters for this iteration.
template <bool FROM_POSE>
__global__ void __launch_bounds__(kThreads)
k_project(const float* __restrict__ verts, const float* __restrict__ Rmat, const float* __restrict__ trans,
          const float* __restrict__ scale, const float* __restrict__ K, float orig, float4* __restrict__ proj,
          int V, int32_t* __restrict__ bin_count, int nstrips, int32_t* __restrict__ loss_counts,
          uint32_t* __restrict__ owned, int owned_words, float* __restrict__ offscreen = nullptr,
          float lw_offscreen = 0.0f, float far_ = 0.0f) {
    const int b = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < owned_words; i += gridDim.x * blockDim.x)
        owned[(size_t)b * owned_words + i] = 0u;
    if (blockIdx.x == 0) {
        if ((int)threadIdx.x < 2 * nstrips) bin_count[b * nstrips * 2 + threadIdx.x] = 0;
        if (loss_counts != nullptr && threadIdx.x < 4) loss_counts[b * 4 + threadIdx.x] = 0;
    }
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    float c[3];
    if (FROM_POSE) {
        float Rm[9], Tm[3], vv[3];
        for (int i = 0; i < 9; i++) Rm[i] = Rmat[9 * b + i];
        for (int i = 0; i < 3; i++) Tm[i] = trans[3 * b + i];
        for (int i = 0; i < 3; i++) vv[i] = verts[3 * v + i];
        transform_vertex(vv, fabsf(scale[0]), Rm, Tm, c);
    } else {
        const float* src = verts + ((size_t)b * V + v) * 3;
        c[0] = src[0]; c[1] = src[1]; c[2] = src[2];
    }
    float Km[6];
    for (int i = 0; i < 6; i++) Km[i] = K[9 * b + i];
    float u, w;
    project_vertex(c, Km, orig, &u, &w);
    proj[(size_t)b * V + v] = make_float4(u, w, c[2], 0.0f);
    if (FROM_POSE && offscreen != nullptr) {
        // stage-1 off-screen penalty (pose_initializtion.py:119-141): how far the projected vertex sticks out of
        // [-1,1]^2 x (0, far), and its gradient through the projection and the rigid transform.  Zero for any sane
        // pose, so the (float, order-dependent) atomics below are off the common path.
        const float ex = fmaxf(u - 1.0f, 0.0f) + fmaxf(-1.0f - u, 0.0f), ey = fmaxf(w - 1.0f, 0.0f) + fmaxf(-1.0f - w, 0.0f);
        const float ez = fmaxf(-c[2], 0.0f) + fmaxf(c[2] - far_, 0.0f);
        const float loss = ex + ey + ez;
        if (loss > 0.0f) {
            const float gu = lw_offscreen * ((u > 1.0f ? 1.0f : 0.0f) - (u < -1.0f ? 1.0f : 0.0f));
            const float gv = lw_offscreen * ((w > 1.0f ? 1.0f : 0.0f) - (w < -1.0f ? 1.0f : 0.0f));
            const float gz = lw_offscreen * ((c[2] > far_ ? 1.0f : 0.0f) - (c[2] < 0.0f ? 1.0f : 0.0f));
            float gc[3];
            project_vertex_backward(c, Km, orig, gu, gv, gc);
            gc[2] += gz;
            float* o = offscreen + (size_t)b * 16;
            const float s_abs = fabsf(scale[0]);
            const float vo[3] = {verts[3 * v], verts[3 * v + 1], verts[3 * v + 2]};
            for (int j = 0; j < 3; j++) atomicAdd(o + j, gc[j]);
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) atomicAdd(o + 3 + 3 * i + j, (s_abs * vo[i]) * gc[j]);
            atomicAdd(o + 13, loss);
        }
    }
}

// ------------------------------------------------------------------------------------------------ binning
// Each (frame, strip) bin holds two groups: faces in their given winding (fn < F) fill the bin from the front,
// reversed windings (fn >= F) from the back.  For an outward-oriented closed mesh the first group is the near
// side of the object, so rasterising it first lets the depth pre-test skip most of the second group.
// bin_count: [B, nstrips, 2].  Counters are first accumulated per CTA in shared memory, so the global atomics
// are one per (CTA, strip, group) instead of one per face.
constexpr int kMaxStrips = kMaxIS / kSH;

__global__ void __launch_bounds__(kThreads)
k_setup_bin(const float4* __restrict__ proj, const int32_t* __restrict__ faces, int V, int F, int is,
            int nstrips, int32_t* __restrict__ bin_count, int32_t* __restrict__ bins) {
    __shared__ int s_cnt[kMaxStrips][2];
    __shared__ int s_base[kMaxStrips][2];
    const int b = blockIdx.y;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int tid = threadIdx.x;
    if (tid < 2 * kMaxStrips) (&s_cnt[0][0])[tid] = 0;
    __syncthreads();
    // per winding the first two strips of the face are kept in registers (statically indexed: slot 2 w + j);
    // anything beyond goes straight to the global counters
    int e_fn[4], e_strip[4], e_slot[4];
#pragma unroll
    for (int k = 0; k < 4; k++) e_fn[k] = -1;
    if (f < F) {
        const float4* P = proj + (size_t)b * V;
        const float4 a0 = P[faces[3 * f + 0]], a1 = P[faces[3 * f + 1]], a2 = P[faces[3 * f + 2]];
#pragma unroll
        for (int w = 0; w < 2; w++) {
            float x[3], y[3];
            x[0] = w ? a2.x : a0.x; y[0] = w ? a2.y : a0.y;
            x[1] = a1.x;            y[1] = a1.y;
            x[2] = w ? a0.x : a2.x; y[2] = w ? a0.y : a2.y;
            int xl, xh, yl, yh;
            if (!face_bbox(x, y, is, &xl, &xh, &yl, &yh)) continue;
            const int fn = f + w * F;
            const int st0 = yl / kSH, st1 = yh / kSH;
#pragma unroll
            for (int j = 0; j < 2; j++)
                if (st0 + j <= st1) {
                    e_fn[2 * w + j] = fn;
                    e_strip[2 * w + j] = st0 + j;
                    e_slot[2 * w + j] = atomicAdd(&s_cnt[st0 + j][w], 1);
                }
            for (int st = st0 + 2; st <= st1; st++) {
                const int slot = atomicAdd(&bin_count[(b * nstrips + st) * 2 + w], 1);
                int32_t* bin = bins + ((size_t)b * nstrips + st) * (size_t)(2 * F);
                bin[w ? 2 * F - 1 - slot : slot] = fn;
            }
        }
    }
    __syncthreads();
    if (tid < 2 * nstrips) {
        const int st = tid >> 1, w = tid & 1;
        const int c = s_cnt[st][w];
        s_base[st][w] = c ? atomicAdd(&bin_count[(b * nstrips + st) * 2 + w], c) : 0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (e_fn[k] < 0) continue;
        const int w = k >> 1;
        const int slot = s_base[e_strip[k]][w] + e_slot[k];
        int32_t* bin = bins + ((size_t)b * nstrips + e_strip[k]) * (size_t)(2 * F);
        bin[w ? 2 * F - 1 - slot : slot] = e_fn[k];
    }
}

// ------------------------------------------------------------------------------------------------ rasteriser
__device__ __forceinline__ void load_face(const float4* __restrict__ P, const int32_t* __restrict__ faces,
                                          int fn, int F, FaceSetup& fs, int* ids) {
    const int w = fn >= F;
    const int f = w ? fn - F : fn;
    const int i0 = faces[3 * f + 0], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    ids[0] = w ? i2 : i0; ids[1] = i1; ids[2] = w ? i0 : i2;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float4 p = P[ids[k]];
        fs.x[k] = p.x; fs.y[k] = p.y; fs.z[k] = p.z;
    }
}

#if DH_DEFER_DEPTH
// Exact key of face g at pixel (xi, yi), recomputed from global memory: used when a hit meets a deferred z-buffer
// entry (below).  Rare, kept out of line.
__device__ __noinline__ unsigned long long exact_key_of(const float4* __restrict__ P, const int32_t* __restrict__ faces,
                                                        int g, int F, int is, int xi, int yi, float near, float far) {
    FaceSetup fg;
    int ids[3];
    load_face(P, faces, g, F, fg, ids);
    face_inverse(fg, is);
    float zg;
    if (!pixel_depth(fg, xi, yi, near, far, &zg)) return DH_ZKEY_EMPTY;
    return zkey(zg, g);
}
#endif

// One queued hit and the z-buffer update.  ent = slot | x << 5 | local_row << 15.
// setup rows: inv[9], z[3], zlo, zhi.  zlo = nearest vertex depth * (1 - 1e-5) (or -inf when that bound does not
// apply), zhi = farthest vertex depth * (1 + 1e-5) when the face may be DEFERRED, else 0.
//
// The face-index map only needs the depth ORDER.  The depth of a face at a pixel (a clamped, normalised harmonic
// blend of its vertex depths) lies in [zlo, zhi], so:
//   * a hit whose zlo is above the depth field of the cell loses without any arithmetic (the pre-test);
//   * the first hit on an empty pixel is stored "deferred": depth field = zhi, bit 31 of the face word set, no
//     division at all.  For a closed mesh rasterised near side first that is almost every covered pixel: the far
//     side then fails the pre-test against zhi;
//   * only a hit that can neither lose nor defer (overlapping depth intervals: silhouette folds, shared edges)
//     computes its exact depth, resolves a deferred occupant exactly as well, and installs the smaller exact key.
// Every losing face is strictly farther than the entry that beat it and the depth field only ever decreases, so
// the final face number is the exact arg-min (lowest face number on exact ties) whatever the order of the hits.
constexpr uint32_t kDeferBit = 0x80000000u;
__device__ __forceinline__ void raster_hit(uint32_t ent, const float (*setup)[32], const int* fns,
                                           unsigned long long* zbuf, int is, int row0, float near, float far,
                                           const float4* __restrict__ P, const int32_t* __restrict__ faces, int F) {
    const int slot = ent & 31, xi = (ent >> 5) & 1023, yl = ent >> 15;
    unsigned long long* cell = zbuf + yl * is + xi;
    unsigned long long cur = *cell;
    const float zlo = setup[12][slot];
    if (__uint_as_float((uint32_t)(cur >> 32)) < zlo) return;   // an empty cell (NaN bits) never compares below
    const int fn = fns[slot];
#if DH_DEFER_DEPTH
    const float zhi = setup[13][slot];
    if (cur == DH_ZKEY_EMPTY && zhi > 0.0f) {
        const unsigned long long want = ((unsigned long long)__float_as_uint(zhi) << 32) | ((uint32_t)fn | kDeferBit);
        const unsigned long long old = atomicCAS(cell, DH_ZKEY_EMPTY, want);
        if (old == DH_ZKEY_EMPTY) return;
        cur = old;
        if (__uint_as_float((uint32_t)(cur >> 32)) < zlo) return;
    }
#endif
    FaceSetup f;
#pragma unroll
    for (int k = 0; k < 9; k++) f.inv[k] = setup[k][slot];
#pragma unroll
    for (int k = 0; k < 3; k++) f.z[k] = setup[9 + k][slot];
    float zp;
    if (!pixel_depth(f, xi, row0 + yl, near, far, &zp)) return;
    const unsigned long long key = zkey(zp, fn);
#if DH_DEFER_DEPTH
    for (;;) {
        unsigned long long best = key;
        if (cur != DH_ZKEY_EMPTY && ((uint32_t)cur & kDeferBit)) {
            const unsigned long long kg = exact_key_of(P, faces, (int)((uint32_t)cur & ~kDeferBit), F, is, xi,
                                                       row0 + yl, near, far);
            if (kg < best) best = kg;
        } else if (cur <= key) {
            return;
        }
        const unsigned long long old = atomicCAS(cell, cur, best);
        if (old == cur) return;
        cur = old;
    }
#else
    if (key < cur) atomicMin(cell, key);
#endif
}

// Per-warp working set of the rasteriser and the pixel-centre table.  They live in the dynamic shared-memory block
// behind the z-buffer strip and are reached through one pointer per warp: accesses to separately declared static
// __shared__ arrays each re-derive the shared-window address (S2UR / UMOV / UIADD3 / ULEA) inside the hot loops.
struct __align__(16) RasterWarp {
    float setup[14][32];   // inv[9], z[3], zlo, zhi of the batch's faces
    float geo[6][32];      // NDC x[3], y[3]
    uint32_t box[32];      // x_lo | x_hi << 10 | local first row << 20
    int start[32];         // first row-item of every face of the batch
    int fn[32];
    uint32_t queue[64];    // pending (lane slot, x, local row) hits
};

// FUSED: epilogue computes the masked-L2 / IoU integer sums and dL/drend (+ sign bitmaps) for this strip.
// else : epilogue writes the pooled, flipped silhouette `rend` (the renderer's return value).
#ifndef DH_RASTER_MIN_CTAS
#define DH_RASTER_MIN_CTAS 2
#endif
template <bool FUSED>
__global__ void __launch_bounds__(kRasterThreads, DH_RASTER_MIN_CTAS)
k_raster(const dh_sil s, const int8_t* __restrict__ mask_tri, float gcoef, float* __restrict__ rend,
         int32_t* __restrict__ loss_counts) {
    extern __shared__ unsigned long long zbuf[];  // [kSH][is]
    __shared__ __align__(16) uint32_t abits[kSH][kMaxIS / 32];
    __shared__ int red[3][kRasterThreads / 32];
    __shared__ int s_next[2];
    __shared__ uint32_t s_tilez[kMaxIS / 16];
    const int is = raster_size(s);
    const int nstrips = is / kSH;
    const int strip = blockIdx.x, b = blockIdx.y;
    const int row0 = strip * kSH;
    const int tid = threadIdx.x;
    const int owned_words = (2 * s.F + 31) >> 5;
    // the face-owns-a-pixel bitmap is staged in shared memory when it is small (<= 4 KB, F <= 16384); for larger
    // meshes the bits go straight to global memory so that the CTA still fits twice per SM
    const bool owned_smem = owned_words <= kOwnedSmemWords;
    uint32_t* s_owned = reinterpret_cast<uint32_t*>(zbuf + kSH * is);
    RasterWarp* s_rw = reinterpret_cast<RasterWarp*>(s_owned + ((owned_smem ? owned_words : 0) + 3) / 4 * 4);
    float* s_ndc = reinterpret_cast<float*>(s_rw + kRasterWarps);   // NDC coordinate of every pixel centre [is]
    int8_t* s_mask = reinterpret_cast<int8_t*>(s_ndc + is);         // the strip's target-mask cells [cell rows][S]
    RasterWarp& RW = s_rw[threadIdx.x >> 5];
    uint32_t* g_owned = s.owned + (size_t)b * owned_words;
    // both passes' entry counts, requested before the z-buffer clear so that their latency is hidden
    const int2 bin_counts = *reinterpret_cast<const int2*>(s.bin_count + (b * nstrips + strip) * 2);
    // a strip no face reaches (above / below the object: about a fifth of them) skips the z-buffer altogether
    const bool empty_strip = (bin_counts.x | bin_counts.y) == 0;
    if (!empty_strip) {
        ulonglong2* z2 = reinterpret_cast<ulonglong2*>(zbuf);   // 16-byte stores: kSH * is is even, the strip is 16-byte aligned
        for (int i = tid; i < kSH * is / 2; i += kRasterThreads) z2[i] = make_ulonglong2(DH_ZKEY_EMPTY, DH_ZKEY_EMPTY);
        for (int i = tid; i < is; i += kRasterThreads) s_ndc[i] = pix_to_ndc(i, is);
    }
    if (owned_smem)
        for (int i = tid; i < owned_words; i += kRasterThreads) s_owned[i] = 0u;
    if (tid < 2) s_next[tid] = 0;
    if (FUSED && strip == 0 && tid == 0) s.gmax[b] = 2.0f * fabsf(gcoef) * (s.aa ? 0.25f : 1.0f);
    __syncthreads();

    const int32_t* bin = s.bins + ((size_t)b * nstrips + strip) * (size_t)(2 * s.F);
    const float4* P = reinterpret_cast<const float4*>(s.proj) + (size_t)b * s.V;
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // The target-mask cells the epilogue scores (one contiguous S-byte row per cell row of the strip) are copied into
    // shared memory with cp.async: the copies run behind the rasterisation and tie up no registers (round 1 held them
    // in six registers per thread; the loads' DRAM latency -- the masks are long evicted from L2 by the time an
    // iteration comes back to them -- then stalled every warp at kernel start, 7.5 % of the kernel's samples).
    if (FUSED) {
        const int S_ = s.S, rows_ = s.aa ? kSH / 2 : kSH, per_row = S_ >> 4;   // 16-byte pieces per row
        for (int i = tid; i < rows_ * per_row; i += kRasterThreads) {
            const int ly = i / per_row, piece = i - ly * per_row;
            const int yo = s.aa ? ((is - 1 - (row0 + 2 * ly)) >> 1) : (is - 1 - (row0 + ly));
            const int8_t* src = mask_tri + ((size_t)b * S_ + yo) * S_ + (piece << 4);
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_mask + ly * S_ + (piece << 4));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // Two passes (given windings, then reversed windings), warp-synchronous batches of 32 bin entries handed out
    // dynamically.  Per batch: (1) lane = face: set-up; (2) lane = (face, row): analytic x-span of the row, then
    // the exact edge tests pixel by pixel, survivors compacted into a per-warp queue; (3) whenever 32 hits are
    // pending, lane = hit: depth (the expensive IEEE divisions) and z-buffer update at full lane occupancy.
#ifndef DH_PASS_ORDER
#define DH_PASS_ORDER 0
#endif
    for (int pass_i = 0; pass_i < 2 && !empty_strip; pass_i++) {
        const int pass = pass_i ^ DH_PASS_ORDER;
        const int count = pass ? bin_counts.y : bin_counts.x;
#if DH_TILE_Z
        if (pass_i == 1) {
            // Between the passes: farthest recorded depth of every 16-column tile of the strip (an empty pixel
            // makes it +inf).  A second-pass face whose nearest vertex lies behind that bound in every tile its
            // bounding box touches cannot win a pixel -- for a closed mesh that is the whole far side, which then
            // costs a vertex fetch and a bounding box instead of spans, edge tests and per-pixel depth pre-tests.
            __syncthreads();
            for (int t = warp; t < (is >> 4); t += kRasterWarps) {
                uint32_t m = 0u;
#pragma unroll
                for (int j = 0; j < (kSH * 16) / 32; j++) {
                    const int q = j * 32 + lane;   // pixel q of the tile: row q / 16, column q % 16
                    m = max(m, (uint32_t)(zbuf[(q >> 4) * is + (t << 4) + (q & 15)] >> 32));
                }
                m = __reduce_max_sync(0xffffffffu, m);
                if (lane == 0) s_tilez[t] = m;
            }
            __syncthreads();
        }
#endif
        // batch schedule: rounds of one 32-face batch per warp, then the remainder split evenly over the warps (a
        // strip holds only a few hundred faces per pass: whole batches would leave most warps idle in the last
        // round).  The z-buffer minimum does not depend on who rasterises what.
#if DH_RASTER_EVEN
        const int full = (count / (32 * kRasterWarps)) * kRasterWarps;
        const int rem = count - 32 * full;
        const int split = pass_i == 0 ? DH_RASTER_SPLIT : DH_RASTER_SPLIT1;
        const int last = (rem + split * kRasterWarps - 1) / (split * kRasterWarps);
#else
        const int full = count / 32, rem = count - 32 * full, last = 32;
#endif
        const int nbatch = full + (rem ? (rem + last - 1) / last : 0);
        for (;;) {
            int k = 0;
            if (lane == 0) k = atomicAdd(&s_next[pass], 1);
            k = __shfl_sync(0xffffffffu, k, 0);
            if (k >= nbatch) break;
            const int base = k < full ? 32 * k : 32 * full + (k - full) * last;
            const int bsize = k < full ? 32 : min(last, count - base);
            const int e = base + lane;
            int nrows = 0;
            if (lane < bsize) {
                const int fn = pass ? bin[2 * s.F - 1 - e] : bin[e];
                FaceSetup fs;
                int ids[3];
                load_face(P, s.faces, fn, s.F, fs, ids);
                bool live = face_bbox(fs.x, fs.y, is, &fs.x_lo, &fs.x_hi, &fs.y_lo, &fs.y_hi);
#if DH_TILE_Z
                if (live && pass_i == 1) {
                    const float zn = fminf(fs.z[0], fminf(fs.z[1], fs.z[2]));
                    if (zn > 0.0f) {   // same bound as the per-pixel pre-test; positive floats order like their bits
                        const uint32_t zb = __float_as_uint(zn * (1.0f - 1e-5f));
                        bool hidden = true;
                        for (int t = fs.x_lo >> 4; t <= (fs.x_hi >> 4); t++) hidden = hidden && s_tilez[t] < zb;
                        live = !hidden;
                    }
                }
#endif
                if (live) {
                    face_inverse(fs, is);
                    const int r_lo = max(fs.y_lo, row0), r_hi = min(fs.y_hi, row0 + kSH - 1);
                    nrows = max(r_hi - r_lo + 1, 0);
#pragma unroll
                    for (int k = 0; k < 9; k++) RW.setup[k][lane] = fs.inv[k];
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        RW.setup[9 + k][lane] = fs.z[k];
                        RW.geo[k][lane] = fs.x[k];
                        RW.geo[3 + k][lane] = fs.y[k];
                    }
                    const float zmin = fminf(fs.z[0], fminf(fs.z[1], fs.z[2]));
                    const float zmax = fmaxf(fs.z[0], fmaxf(fs.z[1], fs.z[2]));
                    const float zlo = (zmin > 0.0f) ? zmin * (1.0f - 1e-5f) : -3.0e38f;
                    const float zhi = zmax * (1.0f + 1e-5f);
                    // deferrable: depth bounds valid and inside (near, far), barycentric set-up well conditioned (the
                    // exact path leaves a pixel unrecorded when its weights degenerate to NaN)
                    bool defer = zmin > 0.0f && s.near_ < zlo && zhi < s.far_;
#pragma unroll
                    for (int k = 0; k < 3; k++)   // NaN fails every comparison
                        defer = defer && fabsf(fs.inv[3 * k]) <= 1.0e3f && fabsf(fs.inv[3 * k + 1]) <= 1.0e3f &&
                                fabsf(fs.inv[3 * k + 2]) <= 1.0e6f;
                    RW.setup[12][lane] = zlo;
                    RW.setup[13][lane] = defer ? zhi : 0.0f;
                    RW.box[lane] = (uint32_t)fs.x_lo | ((uint32_t)fs.x_hi << 10) | ((uint32_t)(r_lo - row0) << 20);
                    RW.fn[lane] = fn;
                }
            }
            int incl = nrows;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            RW.start[lane] = incl - nrows;
            const int total_rows = __shfl_sync(0xffffffffu, incl, 31);
            __syncwarp();
            int qn = 0;
            for (int it0 = 0; it0 < total_rows; it0 += 32) {
                const int it = it0 + lane;
                int xi = 1, xb = 0, slot = 0, rl = 0;
                FaceSetup fs;
                float yp = 0.0f;
                if (it < total_rows) {
                    // last face whose first row-item is <= it (faces without rows share their successor's start)
                    int lo = 0, hi = 31;
#pragma unroll
                    for (int k = 0; k < 5; k++) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (RW.start[mid] <= it) lo = mid; else hi = mid - 1;
                    }
                    slot = lo;
                    const uint32_t box = RW.box[slot];
                    rl = (int)(box >> 20) + (it - RW.start[slot]);
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        fs.x[k] = RW.geo[k][slot];
                        fs.y[k] = RW.geo[3 + k][slot];
                    }
                    yp = s_ndc[row0 + rl];
                    row_span(fs, yp, is, (int)(box & 1023u), (int)((box >> 10) & 1023u), &xi, &xb);
                }
                while (__any_sync(0xffffffffu, xi <= xb)) {
                    bool hit = false;
                    uint32_t ent = 0;
                    if (xi <= xb) {
                        hit = pixel_inside(fs, s_ndc[xi], yp);
                        ent = (uint32_t)slot | ((uint32_t)xi << 5) | ((uint32_t)rl << 15);
                        xi++;
                    }
                    const uint32_t hits = __ballot_sync(0xffffffffu, hit);
                    if (hit) RW.queue[qn + __popc(hits & lt_mask)] = ent;
                    qn += __popc(hits);
                    __syncwarp();
                    if (qn >= 32) {
                        qn -= 32;
                        raster_hit(RW.queue[qn + lane], RW.setup, RW.fn, zbuf, is, row0, s.near_,
                                   s.far_, P, s.faces, s.F);
                        __syncwarp();
                    }
                }
            }
            if (lane < qn)
                raster_hit(RW.queue[lane], RW.setup, RW.fn, zbuf, is, row0, s.near_, s.far_, P, s.faces,
                           s.F);
            __syncwarp();
        }
        // no barrier between the passes: a warp that runs ahead into the reversed windings only makes the depth
        // pre-test less effective, never wrong
    }
    __syncthreads();

    // ---- epilogue 1: face index map + coverage bitmap (one warp = 32 consecutive pixels of a row)
    const int wpr = is >> 5;
    int32_t* fidx = s.fidx + (size_t)b * is * is + (size_t)row0 * is;
    uint32_t* abits_g = s.alpha_bits + ((size_t)b * is + row0) * wpr;
    const int wpr_sh = 31 - __clz(wpr);
    const bool wpr_pow2 = (wpr & (wpr - 1)) == 0;
    if (empty_strip) {
        int4* f4 = reinterpret_cast<int4*>(fidx);   // (row0 * is) * 4 bytes: 16-byte aligned, is % 32 == 0
        for (int i = tid; i < kSH * is / 4; i += kRasterThreads) f4[i] = make_int4(-1, -1, -1, -1);
        for (int i = tid; i < kSH * wpr; i += kRasterThreads) {
            (&abits[0][0])[(i / wpr) * (kMaxIS / 32) + (i % wpr)] = 0u;
            abits_g[i] = 0u;
        }
    }
    if (!empty_strip && (wpr & 3) == 0) {
        // one warp = 128 consecutive pixels of a row per trip, lane l taking pixels l, l + 32, l + 64, l + 96: four
        // independent 8-byte z-buffer reads in flight, coalesced face-index stores, one ballot per coverage word
        const int groups = wpr >> 2;
        for (int seg = warp; seg < kSH * groups; seg += kRasterWarps) {
            const int r = wpr_pow2 ? (seg >> (wpr_sh - 2)) : seg / groups, g4 = seg - r * groups;
            const int i0 = r * is + (g4 << 7) + lane;
            unsigned long long key[4];
#pragma unroll
            for (int j = 0; j < 4; j++) key[j] = zbuf[i0 + 32 * j];
            uint32_t word[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const bool cov = key[j] != DH_ZKEY_EMPTY;
                const int fn = cov ? (int32_t)((uint32_t)key[j] & ~kDeferBit) : -1;
                fidx[i0 + 32 * j] = fn;
                const int fn_left = __shfl_up_sync(0xffffffffu, fn, 1);
                if (cov && (lane == 0 || fn_left != fn))
                    atomicOr(owned_smem ? &s_owned[fn >> 5] : &g_owned[fn >> 5], 1u << (fn & 31));
                word[j] = __ballot_sync(0xffffffffu, cov);
            }
            if (lane == 0) {
                *reinterpret_cast<uint4*>(&abits[r][g4 << 2]) = make_uint4(word[0], word[1], word[2], word[3]);
                *reinterpret_cast<uint4*>(&abits_g[r * wpr + (g4 << 2)]) = make_uint4(word[0], word[1], word[2], word[3]);
            }
        }
    }
    for (int seg = warp; seg < kSH * wpr && !empty_strip && (wpr & 3) != 0; seg += kRasterWarps) {   // narrow rasters
        const int r = wpr_pow2 ? (seg >> wpr_sh) : seg / wpr, cw = seg - r * wpr;
        const int i = r * is + (cw << 5) + lane;
        const unsigned long long key = zbuf[i];
        const bool cov = key != DH_ZKEY_EMPTY;
        const int fn = cov ? (int32_t)((uint32_t)key & ~kDeferBit) : -1;
        fidx[i] = fn;
        // face-owns-a-pixel bitmap (lets the backward skip faces that are completely hidden); runs of the
        // same face along a row set the bit once
        const int fn_left = __shfl_up_sync(0xffffffffu, fn, 1);
        if (cov && (lane == 0 || fn_left != fn))
            atomicOr(owned_smem ? &s_owned[fn >> 5] : &g_owned[fn >> 5], 1u << (fn & 31));
        const uint32_t word = __ballot_sync(0xffffffffu, cov);
        if (lane == 0) {
            abits[r][cw] = word;
            abits_g[r * wpr + cw] = word;
        }
    }
    __syncthreads();
    if (owned_smem)
        for (int i = tid; i < owned_words; i += kRasterThreads) {
            const uint32_t w = s_owned[i];
            if (w) atomicOr(&g_owned[i], w);
        }
    // ---- epilogue 2: output-resolution cells of this strip (2x2 average pool + vertical flip)
    const int S = s.S;
    const int cell_rows = s.aa ? kSH / 2 : kSH;
    const int wprp = (S + 31) >> 5;
    int sse = 0, inter = 0, uni = 0;
    const int wprp_sh = 31 - __clz(wprp);
    const bool wprp_pow2 = (wprp & (wprp - 1)) == 0;
    if (FUSED) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();   // every thread's mask copies have landed
    }
    auto score_cells = [&](int seg) {   // one warp = 32 consecutive cells of a row
        const int ly = wprp_pow2 ? (seg >> wprp_sh) : seg / wprp, x = ((seg - ly * wprp) << 5) + lane;
        int pop, yo;
        if (s.aa) {
            const int r = 2 * ly;
            const uint32_t w0 = abits[r][x >> 4], w1 = abits[r + 1][x >> 4];
            const int sh = (2 * x) & 31;
            pop = __popc((w0 >> sh) & 3u) + __popc((w1 >> sh) & 3u);
            yo = (is - 1 - (row0 + r)) >> 1;
        } else {
            pop = 4 * (int)((abits[ly][x >> 5] >> (x & 31)) & 1u);
            yo = is - 1 - (row0 + ly);
        }
        const size_t o = ((size_t)b * S + yo) * S + x;
        if (FUSED) {
            const int m = s_mask[ly * S + x];
            const int keep = m >= 0, ref = m > 0;
            const int k = keep ? pop - 4 * ref : 0;  // 4 * (image - ref)
            sse += k * k;
            inter += ref ? pop : 0;
            uni += 4 * ref + (keep ? pop : 0) - (ref ? pop : 0);
            // dL/drend = gcoef * (k / 2) is not materialised: the backward rebuilds it from the coverage bitmap and
            // the two sign bitmaps below (grad_from_pop)
            const uint32_t pw = __ballot_sync(0xffffffffu, k > 0);
            const uint32_t nw = __ballot_sync(0xffffffffu, k < 0);
            if (lane == 0) {
                s.pos_pool[((size_t)b * S + yo) * wprp + (x >> 5)] = pw;
                s.neg_pool[((size_t)b * S + yo) * wprp + (x >> 5)] = nw;
            }
        } else {
            rend[o] = (float)pop * 0.25f;
        }
    };
    for (int seg = warp; seg < cell_rows * wprp; seg += kRasterWarps) score_cells(seg);
    if (FUSED) {
        for (int o = 16; o > 0; o >>= 1) {
            sse += __shfl_xor_sync(0xffffffffu, sse, o);
            inter += __shfl_xor_sync(0xffffffffu, inter, o);
            uni += __shfl_xor_sync(0xffffffffu, uni, o);
        }
        if ((tid & 31) == 0) { red[0][tid >> 5] = sse; red[1][tid >> 5] = inter; red[2][tid >> 5] = uni; }
        __syncthreads();
        if (tid < 3) {
            int t = 0;
            for (int w = 0; w < kRasterThreads / 32; w++) t += red[tid][w];
            if (t) atomicAdd(&loss_counts[b * 4 + tid], t);
        }
    }
}

// sign bitmaps and per-frame max magnitude of a caller-provided dL/drend (API backward)
__global__ void __launch_bounds__(kThreads)
k_grad_signs(const float* __restrict__ g, uint32_t* __restrict__ pos_pool, uint32_t* __restrict__ neg_pool,
             float* __restrict__ gmax, long long ncell, int cells_per_frame, float gscale) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const float v = (i < ncell) ? g[i] : 0.0f;
    const uint32_t pw = __ballot_sync(0xffffffffu, v > 0.0f);
    const uint32_t nw = __ballot_sync(0xffffffffu, v < 0.0f);
    float mx = fabsf(v) * gscale;
    if (!(mx <= 3.0e38f)) mx = 0.0f;  // NaN / inf gradients do not size the fixed point
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && i < ncell) {
        pos_pool[i >> 5] = pw;
        neg_pool[i >> 5] = nw;
        if (mx > 0.0f) atomicMax(reinterpret_cast<int*>(gmax) + (int)(i / cells_per_frame), __float_as_int(mx));
    }
}

// ------------------------------------------------------------------------------------------------ backward
// Warp-cooperative edge-scan backward.  The per-face algorithm of dh_core.h::backward_face is cut into three
// stages with per-warp queues in shared memory between them, so that each stage runs at (nearly) full lanes:
//   items  front-facing (face, winding) pairs, compacted from the chunk's faces;
//   phase 1 (lane = item): walks the 6 (edge, axis) spans of the face and emits one task per scan-line crossing
//           that can contribute: OUT if the line has "wanted but uncovered" pixels beyond the edge (O(1) test
//           against per-line first/last ranges), IN if the pixel just outside the edge is uncovered;
//   phase 2 (lane = task): ownership test on the face-index map, then the pixel loop; the two edge-vertex terms
//           are accumulated in 64-bit fixed point (shared-memory atomics), which makes the sum exact and hence
//           independent of task order: bit-identical results run to run.
//   tail    (lane = item): fixed point -> float, projection / rigid-transform backward, pose accumulators.
#ifndef DH_BWD_THREADS
#define DH_BWD_THREADS 352
#endif
#ifndef DH_BWD_MIN_CTAS
#define DH_BWD_MIN_CTAS 2
#endif
#ifndef DH_OWN_EARLY
#define DH_OWN_EARLY 0   // 1 (list path, DH_FLAT_ENUM): the "face owns the pixel inside the edge" test of an out scan is made
#endif                   // during the enumeration, one step behind its face-index load; tasks that fail (18 %) are never
                         // queued.  Measured: 1.055 vs 0.981 ms -- the load's latency in the enumeration costs more
#ifndef DH_GUIDED
#define DH_GUIDED 1      // 1: batch sizes shrink towards the end of the chunk (guided schedule); 0: DH_EVEN_LAST's
#endif
#ifndef DH_GUIDE_MIN
#define DH_GUIDE_MIN 10  // smallest batch of the guided schedule (6 / 8 / 12 / 16 measured: 1.02 / 1.03 / 1.01 / 1.03 ms)
#endif
#ifndef DH_GUIDE_DIV
#define DH_GUIDE_DIV 1   // next batch = remaining items / (warps * DH_GUIDE_DIV), at most 32
#endif
#ifndef DH_EVEN_LAST
#define DH_EVEN_LAST 1
#endif
#ifndef DH_FIDX_NOALLOC
#define DH_FIDX_NOALLOC 0
#endif
#ifndef DH_FAST_COEF
#define DH_FAST_COEF 1
#endif
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}
// one face-index read of the ownership test: random 4-byte reads of a 1 MB map, no reuse
__device__ __forceinline__ int load_fidx(const int32_t* p) {
#if DH_FIDX_NOALLOC
    int v;
    asm volatile("ld.global.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return *p;
#endif
}
constexpr int kBwdThreads = DH_BWD_THREADS;   // 11 warps x 2 CTAs/SM (80 registers, ~97 KB smem); 9 / 10 / 12 warps measured slower
constexpr int kBwdWarps = kBwdThreads / 32;
constexpr int kTaskCap = 64;
constexpr int kMaxBatches = 96;
#ifndef DH_CHUNK_FACES
#define DH_CHUNK_FACES 1024
#endif
constexpr int kChunkFaces = DH_CHUNK_FACES;   // faces per backward CTA at most (item list: 2 windings x 1024 x u16 = 4 KB)
#ifndef DH_PAIR_CAP
#define DH_PAIR_CAP 16
#endif
constexpr int kPairCap = DH_PAIR_CAP;  // pixels one out-scan task handles before it re-queues its remainder

struct __align__(16) BwdWarp {
    float px[3][32], py[3][32];          // pixel coordinates of the batch's faces, by lane slot
    unsigned long long acc[6][32];       // fixed-point sums of the terms, [vertex * 2 + xy][slot]
    uint2 tq[kTaskCap];                  // .x: slot | edge << 5 | axis << 7 | kind << 8 | d0 << 9 | resume d1 << 19
                                         // .y: d1_cross of the task's crossing (float bits; computed once, in phase 1)
};
#ifndef DH_FLAT_ENUM
#define DH_FLAT_ENUM 1   // 1 (list path): crossings of a batch enumerated one per lane over the batch's compacted span list
#endif                   // 0: every lane walks the six spans of its own face
// list path only: the (edge, axis) spans of the faces of the batch, set up at full lanes before the crossing loop.
#if DH_FLAT_ENUM
// The non-empty spans, compacted in (lane, span) order.  With start = number of crossings (scan lines) of the spans
// before it: info = lane | (edge * 2 + axis) << 5 | (direction > 0) << 8 | (d0_from - start + 2^17) << 9 (a crossing's
// scan line is its number + that offset), start16 = start mod 2^16 (only differences below 2^14 are ever taken).
struct __align__(16) BwdSpans {
    float slope[192];
    uint32_t info[192];
    uint16_t start16[192];
};
#else
// rng = d0_from | d0_to << 10 | (direction > 0) << 20 | empty << 21
struct __align__(16) BwdSpans {
    float slope[6][32];
    uint32_t rng[6][32];
};
#endif

// 64-bit fixed-point accumulation in shared memory as two native 32-bit atomics (a 64-bit shared atomicAdd is a
// compare-and-swap loop): the low words add up modulo 2^32, and every add that wraps carries one into the high
// word.  The number of wraps does not depend on the order of the adds, so the total is exact once all are done.
__device__ __forceinline__ void atomic_add_fixed(unsigned long long* a, long long v) {
    if (v == 0) return;
    uint32_t* p = reinterpret_cast<uint32_t*>(a);
    const uint32_t lo = (uint32_t)(unsigned long long)v, hi = (uint32_t)((unsigned long long)v >> 32);
    const uint32_t old = atomicAdd(p, lo);
    const uint32_t h = hi + (((uint32_t)(old + lo) < lo) ? 1u : 0u);
    if (h) atomicAdd(p + 1, h);
}

// Frame-level maps the backward needs, built once per frame instead of once per backward CTA:
//   negT      column-major bitmap of "uncovered && dL/dpixel < 0" (32x32 bit-block transposes through ballots)
//   row_rng   first / last set pixel of every row and of every column of that bitmap
//   neg_lists (fused path) the same pixels as two compressed line lists -- one entry per pixel, grouped by row
//             (axis 1) and by column (axis 0), ascending along the line -- so that an out scan walks the handful of
//             contributing pixels of its line instead of searching bitmap words.  Per (frame, axis): kNLStart u16
//             line starts (start[is] = total, 0xFFFF = more than kNLCap pixels: the frame takes the bitmap path),
//             then kNLCap u16 entries: position along the line | (4 - covered sub-pixels of the pooled cell) << 10.
// One CTA per frame.
#ifndef DH_LISTS_GLOBAL
#define DH_LISTS_GLOBAL 1   // 1: the backward reads the pixel lists from global memory through L1; 0: staged in smem
#endif
#ifndef DH_ALPHA_GLOBAL
#define DH_ALPHA_GLOBAL 0   // 1 (list path): the backward reads the coverage bitmap from global memory through L1 instead of
#endif                      // staging its 32 KB per CTA: smaller CTAs, more of them per SM
constexpr int kNegThreads = 512;
constexpr int kNLStart = 520;              // >= kMaxIS + 1, keeps the entries 16-byte aligned
constexpr int kNLAxis = DH_LISTS_GLOBAL ? 32768 : 8192;  // u16 per (frame, axis)
constexpr int kNLCap = kNLAxis - kNLStart;
constexpr uint32_t kNLOverflow = 0xFFFFu;

__device__ __forceinline__ int block_exclusive_scan_512(int v, int* s_wsum, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();  // s_wsum may still be read from the previous call
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    int base = 0, tot = 0;
    for (int w = 0; w < kNegThreads / 32; w++) {
        const int t = s_wsum[w];
        if (w < warp) base += t;
        tot += t;
    }
    *total = tot;
    return base + incl - v;
}

// 32x32 bit-matrix transpose across a warp: lane k holds row k (bit i = column i) and gets column k back
// (bit i = row i).  Five butterfly stages of one shuffle each (recursive swap of the off-diagonal quadrants).
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {
    uint32_t m = 0x0000FFFFu;
#pragma unroll
    for (int j = 16; j != 0; j >>= 1) {
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
        if ((lane & j) == 0) x ^= (((x >> j) ^ y) & m) << j;
        else                 x ^= ((y >> j) ^ x) & m;
        m ^= m << (j >> 1);
    }
    return x;
}

// frame_coef != NULL (stage-1 IoU loss, pose_initializtion.py:143-155): the loss 1 - I / (U + 1e-6) of a frame has only
// two gradient values -- -lw / (U + eps) on target pixels, +lw I / (U + eps)^2 elsewhere (times keep) -- known once the
// whole frame is rasterised; they are stored per frame, in the units the backward uses for the masked-L2 loss
// (pixel gradient = 2 * coefficient), together with the frame's gradient bound gmax.
__global__ void __launch_bounds__(kNegThreads)
k_neg_maps(const dh_sil s, int build_lists, int list_cap, int32_t* __restrict__ loss_counts,
           float* __restrict__ frame_coef, float lw_iou) {
    extern __shared__ uint32_t nm_words[];  // [is][wpr + 1] row-major; the column-side CTA transposes it in place
                                            // (rows padded by one word: a thread per line and the block transposes
                                            // stay conflict-free); then the frame's coverage bitmap [is][wpr]
    __shared__ int s_wsum[kNegThreads / 32];
    const int is = raster_size(s), S = s.S;
    const int wpr = is >> 5, wprp = (S + 31) >> 5;
    // grid (B, 2): block y = 1 builds the row side (row ranges, row lists), y = 0 the column side (transposed bitmap,
    // column lists)
    const int b = blockIdx.x, tid = threadIdx.x;
    // gridDim.y == 2: one CTA per side; gridDim.y == 1: one CTA does the row side, transposes, then the column side
    const bool single = gridDim.y == 1;
    const int warp = tid >> 5, lane = tid & 31;
    if (frame_coef != nullptr && blockIdx.y == 0 && tid == 0) {   // (either grid shape)
        const float I = (float)loss_counts[b * 4 + 1] * 0.25f, U = (float)loss_counts[b * 4 + 2] * 0.25f + 0.000001f;
        const float g_neg = lw_iou / U, g_pos = (lw_iou * I / U) / U;
        frame_coef[2 * b + 0] = 0.5f * g_neg;
        frame_coef[2 * b + 1] = 0.5f * g_pos;
        s.gmax[b] = fmaxf(g_neg, g_pos);
    }
    uint32_t* words = nm_words;
    const int wps = wpr + 1;
    uint32_t* wordsT = nm_words;            // after the in-place transpose
    uint32_t* s_alpha = nm_words + is * wps;  // the lists' pixel codes read it (long lines would otherwise wait on
                                              // two global loads per listed pixel)
    const uint32_t* ga = s.alpha_bits + (size_t)b * is * wpr;
    const uint32_t* gn = s.neg_pool + (size_t)b * S * wprp;
    for (int i = tid; i < is * wpr; i += kNegThreads) s_alpha[i] = ga[i];
    __syncthreads();
    for (int i = tid; i < is * wpr; i += kNegThreads) {
        const int r = i / wpr, w = i - r * wpr;
        words[r * wps + w] = neg_row_word(s_alpha, gn, is, s.aa, wpr, wprp, r, w);
    }
    __syncthreads();
    // ranges + lists of one axis from its line-major bitmap W (axis 0: lines are columns = words of the transposed
    // bitmap, axis 1: lines are rows)
    auto side = [&](int axis, const uint32_t* W) {
        int cnt = 0, lo = is, hi = -1;
        if (tid < is) {
            for (int w = 0; w < wpr; w++) {
                const uint32_t bits = W[tid * wps + w];
                if (bits) {
                    if (lo == is) lo = (w << 5) + ctz32(bits);
                    hi = (w << 5) + 31 - __clz((int)bits);
                    cnt += __popc(bits);
                }
            }
            // [b][0..1] rows, [b][2..3] columns: every backward CTA of the frame copies them instead of deriving them
            s.row_rng[((size_t)b * 4 + (axis ? 0 : 2)) * is + tid] = (int16_t)lo;
            s.row_rng[((size_t)b * 4 + (axis ? 1 : 3)) * is + tid] = (int16_t)hi;
        }
        if (!build_lists) return;
        int total;
        const int start = block_exclusive_scan_512(cnt, s_wsum, &total);
        uint16_t* L = s.neg_lists + ((size_t)b * 2 + axis) * kNLAxis;
        const bool over = total > list_cap;
        if (tid < is) L[tid] = (uint16_t)(over ? 0 : start);
        if (tid == 0) {
            L[is] = (uint16_t)(over ? kNLOverflow : (uint32_t)total);
            // the frames' overflow flags side by side (slot 3 of the frame's loss counters, zeroed by k_project): the
            // bitmap kernel behind the list kernel reads them in one pass and normally ends right there
            if (over && loss_counts != nullptr) loss_counts[b * 4 + 3] = 1;
        }
        if (over || tid >= is) return;
        uint16_t* E = L + kNLStart + start;
        for (int w = 0; w < wpr; w++) {
            uint32_t bits = W[tid * wps + w];
            while (bits) {
                const int d1 = (w << 5) + ctz32(bits);
                bits &= bits - 1;
                const int r = axis ? tid : d1, c = axis ? d1 : tid;
                int pop = 0;
                if (s.aa) {
                    const int sh = c & 30;
                    const uint32_t w0 = s_alpha[(r & ~1) * wpr + (c >> 5)], w1 = s_alpha[(r | 1) * wpr + (c >> 5)];
                    pop = __popc((w0 >> sh) & 3u) + __popc((w1 >> sh) & 3u);
                }
                *E++ = (uint16_t)((uint32_t)d1 | ((uint32_t)(4 - pop) << 10));
            }
        }
    };
    if (single || blockIdx.y == 1) side(1, words);
    if (single || blockIdx.y == 0) {
        if (single) __syncthreads();   // the row side has read the row-major words
        // in-place transpose: the 32x32 bit blocks (rb, cb) and (cb, rb), rb <= cb, are transposed through ballots
        // and swapped; every warp owns whole block pairs
        const int npairs = wpr * (wpr + 1) / 2;
        for (int pr = warp; pr < npairs; pr += kNegThreads / 32) {
            int rb = 0, rem = pr;
            while (rem >= wpr - rb) { rem -= wpr - rb; rb++; }
            const int cb = rb + rem;
            const uint32_t wa = words[(32 * rb + lane) * wps + cb], wb = words[(32 * cb + lane) * wps + rb];
            const uint32_t ta = transpose32(wa, lane), tb = transpose32(wb, lane);
            __syncwarp();
            words[(32 * cb + lane) * wps + rb] = ta;
            words[(32 * rb + lane) * wps + cb] = tb;
        }
        __syncthreads();
        uint32_t* gT = s.negT + (size_t)b * is * wpr;
        for (int i = tid; i < is * wpr; i += kNegThreads) gT[i] = wordsT[(i / wpr) * wps + (i % wpr)];
        side(0, wordsT);
    }
}

// dL/d(raster pixel) at (r, c).  Fused path: rebuilt from the staged coverage bitmap exactly as k_raster's
// epilogue computed it (gpool = gcoef * (k * 0.5), k = keep * (pop - 4 ref)), so the pixel loops never wait on
// global memory.  `wanted` tells which sign bitmap selected the pixel: true -> k = pop - 4 (< 0, object missing),
// false -> k = pop (> 0, object where background is wanted).  API path: the caller's gradient map.
__device__ __forceinline__ float grad_from_pop(int pop, bool wanted, float gcoef, float gscale) {
    const int k = wanted ? pop - 4 : pop;
    return (gcoef * ((float)k * 0.5f)) * gscale;
}
template <bool FUSED>
__device__ __forceinline__ float grad_value(const BwdMaps& m, int r, int c, bool wanted, float gcoef) {
    if (!FUSED) return grad_at(m, r, c);
    int pop;
    if (m.aa) {
        const int sh = c & 30;
        const uint32_t w0 = m.alpha[(r & ~1) * m.wpr + (c >> 5)], w1 = m.alpha[(r | 1) * m.wpr + (c >> 5)];
        pop = __popc((w0 >> sh) & 3u) + __popc((w1 >> sh) & 3u);
    } else {
        pop = 4 * (int)((m.alpha[r * m.wpr + (c >> 5)] >> (c & 31)) & 1u);
    }
    return grad_from_pop(pop, wanted, gcoef, m.gscale);
}

// phase 2: one task per lane.  The pixel loop accumulates its two terms in fp32 in pixel order (deterministic for
// a given task, like the reference's per-face loop); only the task totals go through the fixed-point atomics.
// Returns 0, or the task word of the remainder when an out scan stops after kPairCap contributing pixels: long
// scans (a line running along the mismatch band) are cut into pieces so that the 32 lanes of a round do similar
// amounts of work.
// LISTS: the out scan walks the line's compressed pixel list (k_neg_maps) instead of bitmap words; task word
//        slot | edge << 5 | axis << 7 | kind << 8 | d0 << 9 (9 bits) | (resume index within the line's list + 1) << 18.
struct NegLists {   // both axes behind one base pointer (indexing an array of pointers by axis would go to local memory)
    const uint16_t* base;
    __device__ __forceinline__ const uint16_t* start(int axis) const { return base + axis * kNLAxis; }
};
template <bool FUSED, bool LISTS>
__device__ __forceinline__ uint32_t bwd_task(uint32_t t, float d1_cross, BwdWarp& W, const BwdMaps& m,
                                             const NegLists& nl, float eps, float fpscale, float gcoef,
                                             float gcoef_pos, const uint16_t* items, int f0, int nF) {
    const int slot = t & 31, edge = (t >> 5) & 3, axis = (t >> 7) & 1, kind = (t >> 8) & 1;
    const int d0 = LISTS ? (int)((t >> 9) & 511u) : (int)((t >> 9) & 1023u);
    const int resume = LISTS ? (int)(t >> 18) : (int)(t >> 19);
    uint32_t cont = 0;
    const int is = m.is;
    // the crossing itself (d1_cross) comes with the task; only the two end points of the edge along the scan
    // axis are needed again for the out scans, the whole span only for the (rare) in scans
    const int e1 = (edge + 1) % 3;
    Span sp;
    sp.p00 = axis ? W.py[edge][slot] : W.px[edge][slot];
    sp.p10 = axis ? W.py[e1][slot] : W.px[e1][slot];
    if (axis == 0) sp.direction = (sp.p00 < sp.p10) ? -1 : 1;
    else           sp.direction = (sp.p00 < sp.p10) ? 1 : -1;
    int d1_in;
    if (0 < sp.direction) d1_in = f2i_sat(floorf(d1_cross));
    else                  d1_in = f2i_sat(ceilf(d1_cross));
    const int d1_out = d1_in + sp.direction;
    if (kind != 0) {  // in scan: the whole span, read by (dynamic) vertex index straight from shared memory
        const int e2 = (edge + 2) % 3;
        sp.p01 = axis ? W.px[edge][slot] : W.py[edge][slot];
        sp.p11 = axis ? W.px[e1][slot] : W.py[e1][slot];
        sp.p20 = axis ? W.py[e2][slot] : W.px[e2][slot];
        sp.p21 = axis ? W.px[e2][slot] : W.py[e2][slot];
        sp.d0_from = f2i_sat(fmaxf(ceilf(fminf(sp.p00, sp.p10)), 0.0f));
        sp.d0_to = f2i_sat(fminf(fmaxf(sp.p00, sp.p10), (float)(is - 1)));
        sp.slope = (sp.p11 - sp.p01) / (sp.p10 - sp.p00);
    }
    const uint32_t item = items[slot];   // the batch's items: local face | winding << 15
    const int fn = f0 + (int)(item & 0x7FFFu) + ((item >> 15) ? nF : 0);
    EdgeCoef ec;
#if DH_FAST_COEF
    if (LISTS) {  // gradients carry a 1e-3 bar: the approximate divider (2 ulp) is enough for the coefficients
        const float num = sp.p10 - sp.p00;
        ec.ka = (sp.p10 != (float)d0) ? __fdividef(num, sp.p10 - (float)d0) : 0.0f;
        ec.kb = (sp.p00 != (float)d0) ? __fdividef(num, (float)d0 - sp.p00) : 0.0f;
    } else
#endif
    edge_coefs(sp.p00, sp.p10, d0, ec);
    const float two_over_is = 2.0f / (float)is;
    float sa = 0.0f, sb = 0.0f;
    if (kind == 0) {
        const int r_in = (axis == 0) ? d1_in : d0, c_in = (axis == 0) ? d0 : d1_in;
        if (LISTS) {
            // Pixels beyond the edge = a suffix (direction +, walked downwards) or a prefix (direction -, walked
            // upwards) of the line's sorted list.  Along an out scan d1 - d1_cross has the sign of the direction, so
            // the sign of dist = k (d1 - d1_cross) (2/is) -- and with it the sign of eps -- is fixed per task, and
            // a skipped term (k == 0) becomes 1 / inf.  -dL/dpixel = code * dunit; dunit multiplies the task's sums.
#if DH_OWN_EARLY && DH_FLAT_ENUM
            const int own = fn;   // tested before the task was queued
            (void)r_in; (void)c_in;
#else
            const int own = load_fidx(m.fidx + r_in * is + c_in);
#endif
            const uint16_t* L = nl.start(axis);
            const int ls = L[d0], le = L[d0 + 1];
            const uint16_t* E = L + kNLStart;
            const bool up = sp.direction < 0;
            const int step = up ? 1 : -1;
            const int i_end = up ? le : ls - 1;
            const int lim = d1_out * step;
            int i = resume ? ls + resume - 1 : (up ? ls : le - 1);   // resume: 1 + index within the line's list
            const int left = (i_end - i) * step;
            const int i_stop = (left > kPairCap) ? i + step * kPairCap : i_end;
            const float dunit = (gcoef * 0.5f) * m.gscale;
            const float ka2 = ec.ka * two_over_is, kb2 = ec.kb * two_over_is;
            const float dirf = (float)sp.direction;
            const float inf = __int_as_float(0x7f800000);
            const float ea = (ec.ka == 0.0f) ? inf : ((0.0f < ka2 * dirf) ? eps : -eps);
            const float eb = (ec.kb == 0.0f) ? inf : ((0.0f < kb2 * dirf) ? eps : -eps);
            if (own == fn && 0.0f < dunit && (ec.ka != 0.0f || ec.kb != 0.0f)) {
                const uint16_t* pe = E + i;   // running pointer: one 64-bit add per pixel instead of re-indexing
                uint32_t e = (i != i_stop) ? *pe : 0u;
                while (i != i_stop) {
                    // the next entry is fetched one iteration ahead (L1 latency behind the two reciprocals); one
                    // entry past either end of a line is still inside the frame's list block
                    pe += step;
                    const uint32_t e_next = *pe;
                    const int d1 = (int)(e & 1023u);
                    if (lim < d1 * step) { i = i_end; break; }  // past d1_out against the walking direction
                    const float x = (float)d1 - d1_cross;
                    const float cf = (float)(e >> 10);
                    float ra, rb;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(fmaf(ka2, x, ea)));
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(fmaf(kb2, x, eb)));
                    sa = fmaf(cf, ra, sa);
                    sb = fmaf(cf, rb, sb);
                    i += step;
                    e = e_next;
                }
                if (i != i_end) cont = (t & 0x3FFFFu) | ((uint32_t)(i - ls + 1) << 18);  // rest of the line: a later round
                sa *= dunit;
                sb *= dunit;
            }
        } else if (m.fidx[r_in * is + c_in] == fn) {
            int from, to;
            out_scan_range(sp.direction, d1_out, is, &from, &to);
            from = max(max(from, (int)((axis == 0) ? m.col_lo[d0] : m.row_lo[d0])), resume);
            to = min(to, (int)((axis == 0) ? m.col_hi[d0] : m.row_hi[d0]));
            const int w_from = from >> 5, w_to = to >> 5;
            int budget = kPairCap;
            for (int w = w_from; w <= w_to && cont == 0; w++) {
                uint32_t bits = neg_line_word(m, axis, d0, w);
                if (w == w_from) bits &= 0xFFFFFFFFu << (from & 31);
                if (w == w_to) bits &= 0xFFFFFFFFu >> (31 - (to & 31));
                // row scans of the fused anti-aliased path: the two coverage words of the pooled row pair serve
                // all 32 pixels of this word
                uint32_t aw0 = 0, aw1 = 0;
                const bool row_fast = FUSED && axis == 1 && m.aa;
                if (row_fast && bits) {
                    aw0 = m.alpha[(d0 & ~1) * m.wpr + w];
                    aw1 = m.alpha[(d0 | 1) * m.wpr + w];
                }
                while (bits) {
                    const int bpos = ctz32(bits);
                    const int d1 = (w << 5) + bpos;
                    if (budget == 0) {  // hand the rest of the line (from pixel d1 on) to a later round
                        cont = (t & 0x7FFFFu) | ((uint32_t)d1 << 19);
                        break;
                    }
                    budget--;
                    bits &= bits - 1;
                    float g;
                    if (row_fast) {
                        const int sh = bpos & 30;
                        g = grad_from_pop(__popc((aw0 >> sh) & 3u) + __popc((aw1 >> sh) & 3u), true, gcoef, m.gscale);
                    } else {
                        g = (axis == 0) ? grad_value<FUSED>(m, d1, d0, true, gcoef)
                                        : grad_value<FUSED>(m, d0, d1, true, gcoef);
                    }
                    const float diff = (0.0f - 1.0f) * g;
                    if (diff <= 0.0f) continue;
                    sa += edge_term_fast(ec.ka, diff, d1, d1_cross, eps, two_over_is);
                    sb += edge_term_fast(ec.kb, diff, d1, d1_cross, eps, two_over_is);
                }
            }
        }
    } else {
        int from, to;
        in_scan_range(sp, d0, d1_in, is, &from, &to);
        for (int d1 = from; d1 <= to; d1++) {
            const int r = (axis == 0) ? d1 : d0, c = (axis == 0) ? d0 : d1;
            if (!alpha_at(m, r, c)) continue;
            if (!pos_at(m, r, c)) continue;
            if (m.fidx[r * is + c] != fn) continue;
            const float diff = (1.0f - 0.0f) * grad_value<FUSED>(m, r, c, false, gcoef_pos);
            if (diff <= 0.0f) continue;
            sa += edge_term_fast(ec.ka, diff, d1, d1_cross, eps, two_over_is);
            sb += edge_term_fast(ec.kb, diff, d1, d1_cross, eps, two_over_is);
        }
    }
    atomic_add_fixed(&W.acc[edge * 2 + (1 - axis)][slot], __float2ll_rn(sa * fpscale));
    atomic_add_fixed(&W.acc[((edge + 1) % 3) * 2 + (1 - axis)][slot], __float2ll_rn(sb * fpscale));
    return cont;
}

// FUSED: accumulate dL/d(T, R, s) of the frame into partials[b][chunk][16].
// else : scatter dL/d(camera-space vertices) into grad_verts [B,V,3] (float atomics).
// LISTS (fused path only): out scans read the per-line pixel lists of k_neg_maps; frames whose lists overflowed
//        are left to the bitmap kernel, which is launched behind it with only_overflow = 1.
template <bool FUSED, bool LISTS>
__device__ __forceinline__ void
bwd_frame(const dh_sil& s, const float* __restrict__ verts_src, const float* __restrict__ Rmat,
          const float* __restrict__ trans, const float* __restrict__ scale, float* __restrict__ partials,
          float* __restrict__ grad_verts, int nchunks, float gcoef_all, const float* __restrict__ frame_coef,
          const int b) {
    extern __shared__ __align__(16) uint32_t smw[];
    __shared__ __align__(16) int16_t s_rng[4][kMaxIS];       // row_lo, row_hi, col_lo, col_hi
    __shared__ uint16_t s_items[2 * kChunkFaces];  // local face | winding << 15, compacted, in face order
    __shared__ float s_bsum[kMaxBatches][13];  // pose-gradient partial sums, one row per batch
#if DH_GUIDED
    __shared__ uint16_t s_bstart[kMaxBatches + 1];
    __shared__ int s_nbatches;
#endif
    __shared__ int s_wcount[kBwdWarps], s_woff[kBwdWarps + 1];
    __shared__ int s_next_batch;
    const int is = raster_size(s), S = s.S;
    const int wpr = is >> 5, wprp = (S + 31) >> 5;
    BwdWarp* s_warp = reinterpret_cast<BwdWarp*>(smw);   // per-warp queues first (static smem is capped at 48 KB)
    BwdSpans* s_spans = reinterpret_cast<BwdSpans*>(smw + kBwdWarps * (sizeof(BwdWarp) / sizeof(uint32_t)));
    uint32_t* s_alpha_smem = smw + kBwdWarps * ((sizeof(BwdWarp) + (LISTS ? sizeof(BwdSpans) : 0)) / sizeof(uint32_t));
    const uint32_t* s_alpha = (LISTS && DH_ALPHA_GLOBAL) ? s.alpha_bits + (size_t)b * is * wpr : s_alpha_smem;
    uint32_t* s_negT = s_alpha_smem + is * wpr;             // bitmap path
    uint32_t* s_negp = s_negT + is * wpr;
#if !DH_LISTS_GLOBAL
    uint16_t* s_lists = reinterpret_cast<uint16_t*>(s_alpha_smem + is * wpr);  // list path: 2 x (starts, entries)
#endif
    const int chunk = blockIdx.x, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint16_t* g_lists = s.neg_lists + (size_t)b * 2 * kNLAxis;
    // ---- stage the frame's maps (built per frame by k_raster / k_neg_maps) with 16-byte asynchronous copies that run
    //      behind the item compaction below: first / last wanted pixel of every row and column, and (list path) the
    //      coverage bitmap
    {
        const int per_row = is / 8;     // uint4 per range row
        for (int i = tid; i < 4 * per_row; i += kBwdThreads) {
            const int k = i / per_row, j = i - k * per_row;
            cp_async16(&s_rng[k][8 * j], s.row_rng + ((size_t)b * 4 + k) * is + 8 * j);
        }
        if (LISTS && !DH_ALPHA_GLOBAL) {
            const uint4* ga4 = reinterpret_cast<const uint4*>(s.alpha_bits + (size_t)b * is * wpr);
            uint4* sa4 = reinterpret_cast<uint4*>(s_alpha_smem);
            for (int i = tid; i < is * wpr / 4; i += kBwdThreads) cp_async16(sa4 + i, ga4 + i);
        }
    }
    const float gmax = s.gmax[b];
    // dL/dpixel coefficients: one number for the masked-L2 loss; per frame (wanted / unwanted pixels) for the IoU loss
    const float gcoef = frame_coef ? frame_coef[2 * b] : gcoef_all;
    const float gcoef_pos = frame_coef ? frame_coef[2 * b + 1] : gcoef_all;

    const float4* P = reinterpret_cast<const float4*>(s.proj) + (size_t)b * s.V;
    const int per = (s.F + nchunks - 1) / nchunks;
    const int f0 = chunk * per, f1 = min(s.F, f0 + per);

    // ---- the chunk's front-facing, pixel-owning (face, winding) items, compacted in face order.  A face that owns
    //      no pixel of the frame contributes nothing: out scans need the in-pixel to be the face's own, in scans
    //      only visit its own pixels.  Each warp scans a contiguous range of faces, <= 8 faces per lane.
    const int per_w = (((f1 - f0 + kBwdWarps - 1) / kBwdWarps) + 31) & ~31;
    const int wf0 = f0 + warp * per_w, wf1 = min(f1, wf0 + per_w);
    uint32_t flags = 0;  // bit 2g: given winding of this lane's g-th face is an item, bit 2g+1: reversed winding
    int wcount = 0;
    if (gmax > 0.0f) {
        const uint32_t* ow = s.owned + (size_t)b * ((2 * s.F + 31) >> 5);
        for (int g = 0, f = wf0 + lane; f - lane < wf1; g++, f += 32) {
            bool v0 = false, v1 = false;
            if (f < wf1) {
                const bool o0 = (ow[f >> 5] >> (f & 31)) & 1u, o1 = (ow[(f + s.F) >> 5] >> ((f + s.F) & 31)) & 1u;
                if (o0 || o1) {
                    const float4 a0 = P[s.faces[3 * f + 0]], a1 = P[s.faces[3 * f + 1]], a2 = P[s.faces[3 * f + 2]];
                    if (finite3(a0.x, a1.x, a2.x) && finite3(a0.y, a1.y, a2.y)) {
                        v0 = o0 && !face_backside(a0.x, a0.y, a1.x, a1.y, a2.x, a2.y);
                        v1 = o1 && !face_backside(a2.x, a2.y, a1.x, a1.y, a0.x, a0.y);
                    }
                }
            }
            flags |= ((uint32_t)v0 << (2 * g)) | ((uint32_t)v1 << (2 * g + 1));
            wcount += __popc(__ballot_sync(0xffffffffu, v0)) + __popc(__ballot_sync(0xffffffffu, v1));
        }
    }
    if (lane == 0) s_wcount[warp] = wcount;
    if (tid == 0) s_next_batch = 0;

    NegLists nl;
    if (LISTS) {
#if DH_LISTS_GLOBAL
        nl.base = g_lists;   // the lists stay in global memory (L1-cached reads)
#else
        const int n_ent = g_lists[is];
        const int n16 = (kNLStart + n_ent + 7) >> 3;
#pragma unroll
        for (int axis = 0; axis < 2; axis++) {
            const uint4* g4 = reinterpret_cast<const uint4*>(g_lists + axis * kNLAxis);
            uint4* s4 = reinterpret_cast<uint4*>(s_lists + axis * kNLAxis);
            for (int i = tid; i < n16; i += kBwdThreads) s4[i] = g4[i];
        }
        nl.base = s_lists;
#endif
    } else {
        const uint32_t* ga = s.alpha_bits + (size_t)b * is * wpr;
        const uint32_t* gt = s.negT + (size_t)b * is * wpr;
        const uint32_t* gn = s.neg_pool + (size_t)b * S * wprp;
        for (int i = tid; i < is * wpr; i += kBwdThreads) { s_alpha_smem[i] = ga[i]; s_negT[i] = gt[i]; }
        for (int i = tid; i < S * wprp; i += kBwdThreads) s_negp[i] = gn[i];
        nl.base = nullptr;
    }
    __syncthreads();
    if (tid == 0) {
        int o = 0;
        for (int w = 0; w < kBwdWarps; w++) { s_woff[w] = o; o += s_wcount[w]; }
        s_woff[kBwdWarps] = o;
#if DH_GUIDED
        // (with one crossing per lane a small batch costs no lanes in the loops, only in its set-up and tail)
        const int nb = bwd_guided_schedule(o, kBwdWarps, DH_GUIDE_MIN, DH_GUIDE_DIV, kMaxBatches, s_bstart);
        s_nbatches = nb;
#endif
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    if (gmax > 0.0f) {
        int pos = s_woff[warp];
        for (int g = 0, f = wf0 + lane; f - lane < wf1; g++, f += 32) {
            const bool v0 = (flags >> (2 * g)) & 1u, v1 = (flags >> (2 * g + 1)) & 1u;
            const uint32_t m0 = __ballot_sync(0xffffffffu, v0), m1 = __ballot_sync(0xffffffffu, v1);
            if (v0) s_items[pos + __popc(m0 & lt_mask)] = (uint16_t)(f - f0);
            if (v1) s_items[pos + __popc(m0) + __popc(m1 & lt_mask)] = (uint16_t)((f - f0) | 0x8000);
            pos += __popc(m0) + __popc(m1);
        }
    }
    __syncthreads();
    const int n_items = s_woff[kBwdWarps];
    (void)n_items;

    BwdMaps m;
    m.alpha = s_alpha; m.neg = nullptr; m.negT = LISTS ? nullptr : s_negT; m.neg_pool = LISTS ? nullptr : s_negp;
    m.pos_pool = s.pos_pool + (size_t)b * S * wprp;   // only the (rare, short) in scans read it: left in global
    m.row_lo = s_rng[0]; m.row_hi = s_rng[1]; m.col_lo = s_rng[2]; m.col_hi = s_rng[3];
    m.gpool = s.gpool + (size_t)b * S * S;
    m.fidx = s.fidx + (size_t)b * is * is;
    m.is = is; m.S = S; m.aa = s.aa; m.wpr = wpr; m.wpr_pool = wprp;
    m.gscale = s.aa ? 0.25f : 1.0f;

    // fixed-point scale: |term| <= gmax / eps; 2^38 / that bound leaves 2^24 terms of headroom in 63 bits
    float fpscale = 0.0f, fpinv = 0.0f;
    if (gmax > 0.0f) {
        int e;
        frexpf(gmax / fmaxf(s.eps, 1e-7f), &e);
        fpscale = ldexpf(1.0f, 38 - e);
        fpinv = ldexpf(1.0f, e - 38);
    }

    float Rm[9], Tm[3], Km[6], s_abs = 1.0f;
    if (FUSED) {
        for (int i = 0; i < 9; i++) Rm[i] = Rmat[9 * b + i];
        for (int i = 0; i < 3; i++) Tm[i] = trans[3 * b + i];
        s_abs = fabsf(scale[0]);
    }
    for (int i = 0; i < 6; i++) Km[i] = s.K[9 * b + i];
    BwdWarp& W = s_warp[warp];
    BwdSpans& SP = s_spans[LISTS ? warp : 0];   // only the list path has (and touches) it
    // Batches of 32 items are handed to the warps dynamically (balance), yet the result does not depend on who
    // ran what: inside a batch every reduction has a fixed order, and each batch leaves its 13 pose-gradient sums
    // in its own row of s_bsum, which are added up in batch order at the end (bit-reproducible results).
    // Batch schedule (a function of n_items only): rounds of kBwdWarps full batches, then the remainder split
    // evenly into kBwdWarps smaller batches, so that the last round keeps every warp busy instead of leaving most
    // of them waiting at the final barrier.
#if DH_GUIDED
    const int n_batches = s_nbatches;
#else
#if DH_EVEN_LAST
    const int full_batches = (n_items / (32 * kBwdWarps)) * kBwdWarps;
#else
    const int full_batches = n_items / 32;
#endif
    const int rem_items = n_items - 32 * full_batches;
    const int last_size = DH_EVEN_LAST ? (rem_items + kBwdWarps - 1) / kBwdWarps : 32;
    const int n_batches = full_batches + (rem_items ? (rem_items + last_size - 1) / last_size : 0);
#endif
    for (;;) {
        int batch = 0;
        if (lane == 0) batch = atomicAdd(&s_next_batch, 1);
        batch = __shfl_sync(0xffffffffu, batch, 0);
        if (batch >= n_batches) break;
#if DH_GUIDED
        const int bstart = s_bstart[batch], bsize = s_bstart[batch + 1] - bstart;
#else
        const int bstart = batch < full_batches ? 32 * batch : 32 * full_batches + (batch - full_batches) * last_size;
        const int bsize = batch < full_batches ? 32 : min(last_size, n_items - bstart);
#endif
        float acc[13];
#pragma unroll
        for (int i = 0; i < 13; i++) acc[i] = 0.0f;
        const bool have = lane < bsize;
        float px[3], py[3];
        int ids[3] = {0, 0, 0};
        if (have) {
            const uint32_t it = s_items[bstart + lane];
            const int fn = f0 + (int)(it & 0x7FFFu) + ((it >> 15) ? s.F : 0);
            FaceSetup fs;
            load_face(P, s.faces, fn, s.F, fs, ids);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                px[k] = ndc_to_pix(fs.x[k], is);
                py[k] = ndc_to_pix(fs.y[k], is);
                W.px[k][lane] = px[k];
                W.py[k][lane] = py[k];
            }
        }
#pragma unroll
        for (int k = 0; k < 6; k++) W.acc[k][lane] = 0ull;
        __syncwarp();
        // ---- phase 1 / 2: enumerate crossings, queue tasks, run them 32 at a time.  ONE loop with ONE inlined copy of
        //      bwd_task (the kernel's hot code stays within the instruction cache): every trip enumerates one step of
        //      crossings and queues its out scans -- or queues the (rare) in scans held back from the step before, so
        //      that the queue never holds more than 31 + 32 tasks -- and then runs full rounds; the last trip drains.
        int n_tasks = 0;
        bool pend_in = false;     // this lane holds an in-scan task back
        uint32_t tw_in = 0;
        float tc_in = 0.0f;
        int stage = 0;            // 1: the next trip queues the held-back in scans (warp-uniform)
        bool more;                // crossings left to enumerate (warp-uniform)
        // DH_OWN_EARLY: out-scan candidates of the previous step, waiting for their face-index load
        bool pq = false, pend_out = false;
        uint32_t ptw = 0;
        float ptc = 0.0f;
        int pown = 0, pfn = 0;
        (void)pq; (void)ptw; (void)ptc; (void)pown; (void)pfn;
#if DH_FLAT_ENUM
        // list path: one crossing per lane.  Every lane sets up the six spans of its face and appends the non-empty
        // ones to the batch's span list (one warp scan gives list positions and crossing offsets); the crossings of the
        // batch are then numbered 0 .. T-1 along that list and handed out 32 at a time, whatever face they belong to:
        // all lanes work until the batch is done, and a lane never switches spans in the middle of the loop.
        int T = 0, NS = 0, s0 = 0, base = 0;   // s0: a span that starts at or before the step's first crossing
        const float* Pb = &W.px[0][0];         // px[3][32] then py[3][32]: [axis][vertex][slot]
        if (LISTS) {
            uint32_t len[6], inf[6], pack = 0;
            float slp[6];
#pragma unroll
            for (int k = 0; k < 6; k++) {
                len[k] = 0; inf[k] = 0; slp[k] = 0.0f;
                if (have) {
                    Span t;
                    span_setup(px, py, k >> 1, k & 1, is, t);
                    const int l = t.d0_to - t.d0_from + 1;
                    len[k] = l > 0 ? (uint32_t)l : 0u;
                    slp[k] = t.slope;
                    inf[k] = span_info(lane, k, t.d0_from, 0 < t.direction, 0u);   // (its start is subtracted below)
                    pack += len[k] + (len[k] ? (1u << 20) : 0u);   // crossings (< 2^17 per batch) | spans << 20
                }
            }
            uint32_t incl = pack;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t totals = __shfl_sync(0xffffffffu, incl, 31);
            T = (int)(totals & 0xFFFFFu);
            NS = (int)(totals >> 20);
            uint32_t tb = (incl - pack) & 0xFFFFFu, cb = (incl - pack) >> 20;
#pragma unroll
            for (int k = 0; k < 6; k++)
                if (len[k]) {
                    SP.start16[cb] = (uint16_t)tb; SP.slope[cb] = slp[k]; SP.info[cb] = inf[k] - (tb << 9);
                    cb++; tb += len[k];
                }
            __syncwarp();
        }
#endif
        // bitmap path (and DH_FLAT_ENUM = 0): every lane walks the six spans of its own face
        int span_id = have ? 0 : 6, d0 = 0;
        Span sp;
        sp.d0_to = -1;
        if (!(LISTS && DH_FLAT_ENUM)) {
            if (have) {
#if !DH_FLAT_ENUM
                if (LISTS) {
#pragma unroll
                    for (int k = 0; k < 6; k++) {
                        Span t;
                        span_setup(px, py, k >> 1, k & 1, is, t);
                        SP.slope[k][lane] = t.slope;
                        SP.rng[k][lane] = (uint32_t)t.d0_from | ((uint32_t)max(t.d0_to, 0) << 10) |
                                          ((0 < t.direction) ? (1u << 20) : 0u) | ((t.d0_to < t.d0_from) ? (1u << 21) : 0u);
                        if (k == 0) { sp = t; d0 = t.d0_from; }
                    }
                } else
#endif
                {
                    span_setup(px, py, 0, 0, is, sp);
                    d0 = sp.d0_from;
                }
            }
            more = __any_sync(0xffffffffu, span_id < 6);
        } else {
#if DH_FLAT_ENUM
            more = T > 0;
#endif
        }
        for (;;) {
            bool q = false;
            uint32_t tw = 0;
            float tc = 0.0f;
            if (stage) {
                q = pend_in; tw = tw_in; tc = tc_in;
                stage = 0;
            } else if (more || pend_out) {
                bool t_out = false, t_in = false;
                int own_new = -1, fn_new = 0;
                (void)own_new; (void)fn_new;
                // (axis, scan line, crossing, pixel inside / outside) -> which scans can contribute.  Out scan: iff the
                // line has a wanted pixel at or beyond d1_out in the scan direction -- one compare against the line's
                // last (direction +) or first (direction -) one.  In scan: iff the pixel outside is uncovered.
                auto classify = [&](int slot, int edge, int axis, int d0_, bool dpos, int d1_out, float d1_cross) {
                    const int bound = s_rng[(axis ? 0 : 2) + (dpos ? 1 : 0)][d0_];
                    t_out = dpos ? (d1_out <= bound) : (bound <= d1_out);
#if DH_OWN_EARLY && DH_FLAT_ENUM
                    if (LISTS && t_out) {   // the load is consumed one trip later
                        const int d1_in = d1_out - (dpos ? 1 : -1);
                        const int r_in = (axis == 0) ? d1_in : d0_, c_in = (axis == 0) ? d0_ : d1_in;
                        own_new = load_fidx(m.fidx + r_in * is + c_in);
                        const uint32_t item = s_items[bstart + slot];
                        fn_new = f0 + (int)(item & 0x7FFFu) + ((item >> 15) ? s.F : 0);
                    }
#endif
                    const int r_out = (axis == 0) ? d1_out : d0_, c_out = (axis == 0) ? d0_ : d1_out;
                    t_in = !((s_alpha[r_out * wpr + (c_out >> 5)] >> (c_out & 31)) & 1u);
                    tw = (uint32_t)slot | ((uint32_t)edge << 5) | ((uint32_t)axis << 7) | ((uint32_t)d0_ << 9);
                    tc = d1_cross;
                };
#if DH_FLAT_ENUM
                if (LISTS && more) {
                    // spans that start inside this step mark their first crossing; a lane's span = s0 + marks up to itself
                    uint32_t bit = 0;
                    const int j = s0 + 1 + lane;
                    if (j < NS) bit = span_mark(SP.start16[j], base);
                    const uint32_t M = __reduce_or_sync(0xffffffffu, bit);
                    const int span = span_of_lane(s0, M, lane);
                    s0 += __popc(M);
                    const int idx = base + lane;
                    if (idx < T) {
                        const uint32_t info = SP.info[span];
                        const int slot = (int)(info & 31u), k = (int)((info >> 5) & 7u), axis = k & 1, edge = k >> 1;
                        const int d0_ = span_info_d0(info, idx);
                        const bool dpos = (info >> 8) & 1u;
                        const float p00 = Pb[(axis * 3 + edge) * 32 + slot], p01 = Pb[((1 - axis) * 3 + edge) * 32 + slot];
                        float c = SP.slope[span] * ((float)d0_ - p00);   // span_crossing (dh_core.h), same operation order
                        c = c + p01;
                        const int d1_in = f2i_sat(dpos ? floorf(c) : ceilf(c));
                        const int d1_out = d1_in + (dpos ? 1 : -1);
                        if ((unsigned)d1_in < (unsigned)is && (unsigned)d1_out < (unsigned)is)
                            classify(slot, edge, axis, d0_, dpos, d1_out, c);
                    }
                    base += 32;
                    more = base < T;
                } else if (!(LISTS && DH_FLAT_ENUM))
#endif
                {
                    if (span_id < 6) {
                        if (d0 <= sp.d0_to) {
                            float d1_cross;
                            int d1_in, d1_out;
                            if (span_crossing(sp, d0, is, &d1_cross, &d1_in, &d1_out))
                                classify(lane, span_id >> 1, span_id & 1, d0, 0 < sp.direction, d1_out, d1_cross);
                            d0++;
                        }
                        if (d0 > sp.d0_to) {  // next span of this lane's face
                            span_id++;
                            if (span_id < 6) {
#if !DH_FLAT_ENUM
                                if (LISTS) {   // (its set-up was done at full lanes above)
                                    const int e0 = span_id >> 1, ax = span_id & 1;
                                    const uint32_t rg = SP.rng[span_id][lane];
                                    sp.slope = SP.slope[span_id][lane];
                                    sp.p00 = ax ? W.py[e0][lane] : W.px[e0][lane];
                                    sp.p01 = ax ? W.px[e0][lane] : W.py[e0][lane];
                                    sp.direction = (rg >> 20) & 1u ? 1 : -1;
                                    d0 = (int)(rg & 1023u);
                                    sp.d0_to = (rg >> 21) & 1u ? -1 : (int)((rg >> 10) & 1023u);
                                } else
#endif
                                {
                                    span_setup(px, py, span_id >> 1, span_id & 1, is, sp);
                                    d0 = sp.d0_from;
                                }
                            }
                        }
                    }
                    more = __any_sync(0xffffffffu, span_id < 6);
                }
                pend_in = t_in; tw_in = tw | (1u << 8); tc_in = tc;
                stage = __any_sync(0xffffffffu, t_in) ? 1 : 0;
#if DH_OWN_EARLY && DH_FLAT_ENUM
                if (LISTS) {
                    // queue the candidates of the step before (their owner has arrived), hold this step's back
                    q = pq && pown == pfn;
                    const uint32_t tw_q = ptw;
                    const float tc_q = ptc;
                    pq = t_out; ptw = tw; ptc = tc; pown = own_new; pfn = fn_new;
                    pend_out = __any_sync(0xffffffffu, t_out);
                    tw = tw_q; tc = tc_q;
                } else
#endif
                    q = t_out;
            }
            const uint32_t mq = __ballot_sync(0xffffffffu, q);
            if (q) W.tq[n_tasks + __popc(mq & lt_mask)] = make_uint2(tw, __float_as_uint(tc));
            n_tasks += __popc(mq);
            __syncwarp();
            const bool last = !more && !stage && !pend_out;
            while (n_tasks >= 32 || (last && n_tasks > 0)) {   // (the drain's remainders are drained too)
                const int nt = min(n_tasks, 32);
                n_tasks -= nt;
                uint32_t cont = 0;
                float dc = 0.0f;
                if (lane < nt) {
                    const uint2 tk = W.tq[n_tasks + lane];
                    dc = __uint_as_float(tk.y);
                    cont = bwd_task<FUSED, LISTS>(tk.x, dc, W, m, nl, s.eps, fpscale, gcoef, gcoef_pos, s_items + bstart, f0,
                                                  s.F);
                }
                __syncwarp();
                const uint32_t mc = __ballot_sync(0xffffffffu, cont != 0u);
                if (cont) W.tq[n_tasks + __popc(mc & lt_mask)] = make_uint2(cont, __float_as_uint(dc));
                n_tasks += __popc(mc);
                __syncwarp();
            }
            if (last) break;
        }
        // ---- tail: per item, fixed point -> float, then through the projection and the rigid transform
        if (have) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float gu = -(float)((double)(long long)W.acc[2 * k][lane] * (double)fpinv);
                const float gv = -(float)((double)(long long)W.acc[2 * k + 1][lane] * (double)fpinv);
                if (gu == 0.0f && gv == 0.0f) continue;
                float c[3], vo[3], gc[3];
                if (FUSED) {
                    for (int i = 0; i < 3; i++) vo[i] = verts_src[3 * ids[k] + i];
                    transform_vertex(vo, s_abs, Rm, Tm, c);
                } else {
                    const float* src = verts_src + ((size_t)b * s.V + ids[k]) * 3;
                    c[0] = src[0]; c[1] = src[1]; c[2] = src[2];
                }
                project_vertex_backward(c, Km, s.orig_size, gu, gv, gc);
                if (FUSED) {
                    for (int j = 0; j < 3; j++) acc[j] += gc[j];
                    for (int i = 0; i < 3; i++)
                        for (int j = 0; j < 3; j++) acc[3 + 3 * i + j] += (s_abs * vo[i]) * gc[j];
                    float dot = 0.0f;
                    for (int j = 0; j < 3; j++)
                        dot += (vo[0] * Rm[j] + vo[1] * Rm[3 + j] + vo[2] * Rm[6 + j]) * gc[j];
                    acc[12] += dot;
                } else {
                    float* dst = grad_verts + ((size_t)b * s.V + ids[k]) * 3;
                    atomicAdd(dst + 0, gc[0]);
                    atomicAdd(dst + 1, gc[1]);
                    atomicAdd(dst + 2, gc[2]);
                }
            }
        }
        if (FUSED) {
#pragma unroll
            for (int i = 0; i < 13; i++) {
                for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
                if (lane == 0) s_bsum[batch][i] = acc[i];
            }
        }
        __syncwarp();
    }
    if (FUSED) {
        __syncthreads();
        if (tid < 16) {
            float t = 0.0f;
            if (tid < 13)
                for (int k = 0; k < n_batches; k++) t += s_bsum[k][tid];
            partials[((size_t)b * nchunks + chunk) * 16 + tid] = t;
        }
    }
}

// Grid (chunks, frames).  only_overflow (the bitmap kernel launched behind the list kernel): a small grid whose rows
// walk the frames and work only on those whose lists overflowed -- normally none, and the launch ends after B / gridDim.y
// flag reads per CTA instead of starting B x chunks CTAs that return at once.
template <bool FUSED, bool LISTS>
__global__ void __launch_bounds__(kBwdThreads, DH_BWD_MIN_CTAS)
k_backward(const dh_sil s, const float* __restrict__ verts_src, const float* __restrict__ Rmat,
           const float* __restrict__ trans, const float* __restrict__ scale, float* __restrict__ partials,
           float* __restrict__ grad_verts, int nchunks, float gcoef_all, int only_overflow,
           const float* __restrict__ frame_coef, const int32_t* __restrict__ loss_counts) {
    const int is = raster_size(s);
    if (LISTS) {
        const int b = blockIdx.y;
        if (s.neg_lists[(size_t)b * 2 * kNLAxis + is] == kNLOverflow) return;
        bwd_frame<FUSED, LISTS>(s, verts_src, Rmat, trans, scale, partials, grad_verts, nchunks, gcoef_all, frame_coef, b);
    } else if (only_overflow) {
        int any = 0;
        for (int b = threadIdx.x; b < s.B; b += kBwdThreads) any |= loss_counts[b * 4 + 3];
        if (!__syncthreads_or(any)) return;
        for (int b = blockIdx.y; b < s.B; b += gridDim.y) {
            if (loss_counts[b * 4 + 3] == 0) continue;   // (uniform over the CTA)
            bwd_frame<FUSED, LISTS>(s, verts_src, Rmat, trans, scale, partials, grad_verts, nchunks, gcoef_all, frame_coef,
                                    b);
            __syncthreads();   // the next frame reuses the shared arrays
        }
    } else {
        bwd_frame<FUSED, LISTS>(s, verts_src, Rmat, trans, scale, partials, grad_verts, nchunks, gcoef_all, frame_coef,
                                blockIdx.y);
    }
}

// ------------------------------------------------------------------------------------------------ pose kernels
// ---- peer-to-peer mailboxes (layout: include/dynhor_b200.h)
__device__ __forceinline__ float* mb_slot(float* mb, int side, int tick) { return mb + (side * 4 + (tick & 3)) * 16; }
__device__ __forceinline__ volatile int* mb_flag(float* mb, int side, int tick) {
    return reinterpret_cast<volatile int*>(mb + 128) + side * 4 + (tick & 3);
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Spin until *flag has reached `tick` (wrap-safe).  Wall-clock bounded: after timeout_ms (default 60 s) -- or at once
// when an earlier wait of this run already failed -- DH_STATUS_HALO_TIMEOUT goes into *status and the caller carries
// on with whatever the slot holds, so that the kernel (and the graph) always terminates; the host raises after
// the run.  A neighbour that is merely slow (lazy peer mapping, graph instantiation, a debugger) is waited for.
__device__ bool mb_wait(volatile int* flag, int tick, int timeout_ms, int32_t* status) {
    if (status != nullptr && *reinterpret_cast<volatile int32_t*>(status) != 0) return false;
    const unsigned long long t0 = globaltimer_ns();
    const unsigned long long limit = (unsigned long long)(timeout_ms > 0 ? timeout_ms : 60000) * 1000000ull;
    while (*flag - tick < 0) {
        if (globaltimer_ns() - t0 > limit) {
            if (status != nullptr) atomicExch(status, DH_STATUS_HALO_TIMEOUT);
            return false;
        }
        __nanosleep(128);
    }
    __threadfence_system();
    return true;
}
// wait until the neighbour has published the pose valid for `tick`, then copy it out (9 floats)
__device__ void mb_wait_read(const dh_jointopt& p, int side, int tick, float* out9) {
    mb_wait(mb_flag(p.mailbox, side, tick), tick, p.halo_timeout_ms, p.status);
    const volatile float* src = mb_slot(p.mailbox, side, tick);
    for (int i = 0; i < 9; i++) out9[i] = src[i];
}
// publish this rank's boundary pose for `tick` into a neighbour's mailbox
__device__ void mb_publish(float* peer, int side, int tick, const float* rot6d, const float* trans) {
    volatile float* dst = mb_slot(peer, side, tick);
    for (int i = 0; i < 6; i++) dst[i] = rot6d[i];
    for (int i = 0; i < 3; i++) dst[6 + i] = trans[i];
    __threadfence_system();
    *mb_flag(peer, side, tick) = tick;
}
// scale slots: one exact partial gradient per (tick & 3, rank)
__device__ __forceinline__ volatile unsigned long long* mb_scale_slot(float* mb, int tick, int rank) {
    return reinterpret_cast<volatile unsigned long long*>(mb + 136 + ((tick & 3) * DH_MAX_RANKS + rank) * 4);
}
__device__ __forceinline__ volatile int* mb_scale_flag(float* mb, int tick, int rank) {
    return reinterpret_cast<volatile int*>(mb + 392) + (tick & 3) * DH_MAX_RANKS + rank;
}

__global__ void k_pose_prep(const dh_jointopt p) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int B = p.sil.B;
    if (b >= B) return;
    float r6[6], Rm[9];
    for (int i = 0; i < 6; i++) r6[i] = p.rot6d[6 * b + i];
    rot6d_to_R(r6, Rm);
    for (int i = 0; i < 9; i++) p.Rmat[9 * b + i] = Rm[i];
    if (p.offscreen != nullptr)
        for (int i = 0; i < 16; i++) p.offscreen[(size_t)b * 16 + i] = 0.0f;
    const float* hp = p.halo_prev;
    const float* hn = p.halo_next;
    float hbuf[2][9];
    if (p.mailbox != nullptr) {  // P2P mode: the neighbours' poses arrive in the mailbox
        const int tick = p.tick_base + *p.step;
        hp = hn = nullptr;
        if (p.peer_prev != nullptr) {
            if (b == 0) mb_wait_read(p, 0, tick, hbuf[0]);
            hp = hbuf[0];
        }
        if (p.peer_next != nullptr) {
            if (b == B - 1) mb_wait_read(p, 1, tick, hbuf[1]);
            hn = hbuf[1];
        }
    }
    // iteration clock (optional): the moment this rank's own work of the iteration can start = after its waits
    if (p.iter_ns != nullptr && (b == 0 || b == B - 1) && *p.step < p.max_iters)
        atomicMax(p.iter_ns + 2 * (size_t)*p.step, globaltimer_ns());
    smooth_terms_frame(b, B, p.rot6d, p.trans, hp, hn, p.scale[0], p.moments, p.sil.V, p.B_total, p.lw_smooth,
                       p.smooth_terms + (size_t)b * 16);
}

// mode 0: Adam update in place.  mode 1: write gradients to (grad_rot6d, grad_trans), leave parameters alone.
// mode 2: per-frame loss terms only (forward-only evaluation; no backward ran, partials are not read).
// One WARP per frame: the frame's partial-sum rows (16 floats each: one per backward chunk, one per correspondence
// segment) are read by lane = (row parity, component) with all loads of a lane independent, added up in double
// (even rows, odd rows, then the two halves), and handed to lane 0, which does the rest.  (A thread per frame read its
// 12 rows one after the other behind a run-time trip count: 25 us for 300 frames.)
constexpr int kPoseUpdateThreads = 128;
__device__ __forceinline__ double shfl_double(double v, int src) {
    return __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(v), src),
                            __shfl_sync(0xffffffffu, __double2loint(v), src));
}
__device__ __forceinline__ double warp_row_sums(const float* __restrict__ rows, int nrows, int lane) {
    const int comp = lane & 15, half = lane >> 4;
    double acc = 0.0;
#pragma unroll 4
    for (int c = half; c < nrows; c += 2) acc += (double)rows[c * 16 + comp];
    return acc + shfl_double(acc, (lane + 16) & 31);   // lanes l and l + 16 now both hold component l & 15
}
__global__ void __launch_bounds__(kPoseUpdateThreads)
k_pose_update(const dh_jointopt p, int mode, float* __restrict__ grad_rot6d,
              float* __restrict__ grad_trans, int with_sil, int with_corr) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (kPoseUpdateThreads / 32) + (threadIdx.x >> 5);
    if (b >= p.sil.B) return;
    double sil_c = 0.0, corr_c = 0.0;   // component (lane & 15) of the frame's sums
    if (with_sil && mode != 2) sil_c = warp_row_sums(p.partials + (size_t)b * p.nchunks * 16, p.nchunks, lane);
    if (with_corr) corr_c = warp_row_sums(p.corr.partials + (size_t)b * p.corr.nslots * 16, p.corr.nslots, lane);
    double G[9], gT[3], gs = 0.0, q[13];
#pragma unroll
    for (int i = 0; i < 13; i++) {
        const double v = shfl_double(sil_c, i);
        q[i] = shfl_double(corr_c, i);
        if (i < 3) gT[i] = v;
        else if (i < 12) G[i - 3] = v;
        else gs = v;
    }
    if (lane != 0) return;
    if (with_sil && mode != 2) gs *= (p.scale[0] < 0.0f) ? -1.0 : 1.0;
    const double* st = p.smooth_terms + (size_t)b * 16;
    for (int i = 0; i < 3; i++) gT[i] += st[i];
    for (int i = 0; i < 9; i++) G[i] += st[3 + i];
    gs += st[12];
    // correspondence term (builder-defined, dh_corr.cu): segment sums, then lw / sum(w)
    double corr_loss = 0.0;
    if (with_corr) {
        const double coef = p.corr.lw_corr / (p.corr.w_sum_dev ? *p.corr.w_sum_dev : p.corr.w_sum);
        const double sc = (double)p.scale[0], s_abs = fabs(sc), sgn = (sc < 0.0) ? -1.0 : 1.0;
        double dot = 0.0;
        for (int i = 0; i < 3; i++) gT[i] += coef * q[i];
        for (int i = 0; i < 9; i++) {
            G[i] += coef * s_abs * q[3 + i];
            dot += (double)p.Rmat[9 * b + i] * q[3 + i];
        }
        gs += coef * sgn * dot;
        corr_loss = q[12];
    }
    double off_loss = 0.0;
    if (p.offscreen != nullptr && mode != 2) {   // stage-1 off-screen penalty: accumulated by k_project
        const float* o = p.offscreen + (size_t)b * 16;
        for (int i = 0; i < 3; i++) gT[i] += (double)o[i];
        for (int i = 0; i < 9; i++) G[i] += (double)o[3 + i];
    }
    if (p.offscreen != nullptr) off_loss = (double)p.offscreen[(size_t)b * 16 + 13];
    float r6[6];
    for (int i = 0; i < 6; i++) r6[i] = p.rot6d[6 * b + i];
    double g6[6];
    rot6d_backward(r6, G, g6);

    double* ft = p.frame_terms + (size_t)b * 8;
    const int32_t* lc = p.loss_counts + b * 4;
    ft[0] = with_sil ? (double)lc[0] : 0.0;
    ft[1] = with_sil ? (double)(((float)lc[1] * 0.25f) / ((float)lc[2] * 0.25f + 0.000001f)) : 0.0;
    ft[2] = st[13];
    ft[3] = gs;
    ft[4] = corr_loss;
    ft[5] = off_loss;

    if (mode == 2) return;
    if (mode == 1) {
        for (int i = 0; i < 6; i++) grad_rot6d[6 * b + i] = (float)g6[i];
        for (int i = 0; i < 3; i++) grad_trans[3 * b + i] = (float)gT[i];
        return;
    }
    const int t = *p.step + 1;
    float step_rot, step_tr, bc2s;
    adam_bias(t, p.lr * (p.loss_mode == DH_LOSS_STAGE1 ? 1.0 : 10.0), &step_rot, &bc2s);   // one group in stage 1
    adam_bias(t, p.lr, &step_tr, &bc2s);
    for (int i = 0; i < 6; i++)
        adam_update(&p.rot6d[6 * b + i], &p.adam_m_rot[6 * b + i], &p.adam_v_rot[6 * b + i], (float)g6[i],
                    step_rot, bc2s);
    for (int i = 0; i < 3; i++)
        adam_update(&p.trans[3 * b + i], &p.adam_m_trans[3 * b + i], &p.adam_v_trans[3 * b + i], (float)gT[i],
                    step_tr, bc2s);
    // P2P halo: the updated boundary poses go straight into the neighbours' mailboxes, valid for iteration t
    if (p.mailbox != nullptr) {
        const int tick = p.tick_base + t;
        if (b == 0 && p.peer_prev != nullptr) mb_publish(p.peer_prev, 1, tick, p.rot6d + 6 * b, p.trans + 3 * b);
        if (b == p.sil.B - 1 && p.peer_next != nullptr)
            mb_publish(p.peer_next, 0, tick, p.rot6d + 6 * b, p.trans + 3 * b);
    }
}

// One CTA.  mode 0: history row + scale update + step++.  mode 1: row max_iters only (grad_scale written).
// mode 2: row max_iters only (forward-only evaluation).
// The scale gradient (one number for the whole sequence, jointopt.py:42-46) is summed exactly (Fx128): first over
// this rank's frames, then -- DH_SCALE_P2P -- over the ranks, every rank storing its partial into all mailboxes and
// adding the `world` partials up in rank order, so that all ranks apply the same bits and a sharded run equals the
// single-GPU run.
__global__ void __launch_bounds__(kThreads) k_finalize(const dh_jointopt p, int mode, float* __restrict__ grad_scale) {
    __shared__ double red[kThreads / 32][4];
    __shared__ Fx128 redg[kThreads / 32];
    __shared__ Fx128 s_parts[DH_MAX_RANKS];
    double a[4] = {0.0, 0.0, 0.0, 0.0};   // 16*SSE, iou, pair_sse (stage 1: off-screen loss), corr loss
    Fx128 g;
    g.hi = 0; g.lo = 0ull;
    for (int b = threadIdx.x; b < p.sil.B; b += kThreads) {
        const double* ft = p.frame_terms + (size_t)b * 8;
        a[0] += ft[0]; a[1] += ft[1]; a[2] += (p.loss_mode == DH_LOSS_STAGE1) ? ft[5] : ft[2]; a[3] += ft[4];
        g = fx_add(g, fx_from_double(ft[3]));
    }
    for (int o = 16; o > 0; o >>= 1) {
        for (int i = 0; i < 4; i++) a[i] += __shfl_xor_sync(0xffffffffu, a[i], o);
        Fx128 t;
        t.hi = __shfl_xor_sync(0xffffffffu, g.hi, o);
        t.lo = __shfl_xor_sync(0xffffffffu, g.lo, o);
        g = fx_add(g, t);
    }
    if ((threadIdx.x & 31) == 0) {
        for (int i = 0; i < 4; i++) red[threadIdx.x >> 5][i] = a[i];
        redg[threadIdx.x >> 5] = g;
    }
    __syncthreads();
    const int step = *p.step;
    if (threadIdx.x == 0) {
        double t[4] = {0.0, 0.0, 0.0, 0.0};
        Fx128 tg;
        tg.hi = 0; tg.lo = 0ull;
        for (int w = 0; w < kThreads / 32; w++) {
            for (int i = 0; i < 4; i++) t[i] += red[w][i];
            tg = fx_add(tg, redg[w]);
        }
        const int row = (mode == 0) ? step : p.max_iters;
        if (mode != 0 || step < p.max_iters) {
            double* h = p.hist + (size_t)row * 4;
            const double N = (double)(p.B_total - 1) * (double)p.sil.V * 3.0;
            if (p.loss_mode == DH_LOSS_STAGE1) {   // sums over the frames of (off-screen penalty, 1 - IoU); mean IoU
                h[0] = t[2];
                h[1] = (double)p.sil.B - t[1];
                h[2] = t[1] / (double)p.B_total;
            } else {
                h[0] = (p.B_total > 1) ? t[2] / N : 0.0;
                h[1] = t[0] / 16.0 / p.keep_sum / (double)p.B_total;
                h[2] = t[1] / (double)p.B_total;
            }
            h[3] = (p.corr.records != nullptr && p.corr.lw_corr > 0.0)
                       ? t[3] / (p.corr.w_sum_dev ? *p.corr.w_sum_dev : p.corr.w_sum) : 0.0;
        }
        redg[0] = tg;   // this rank's exact partial
        if (mode == 1 && grad_scale != nullptr) grad_scale[0] = (float)fx_to_double(tg);
    }
    if (mode != 0) return;
    __syncthreads();
    if (p.optimize_scale) {
        const Fx128 mine = redg[0];
        if (p.scale_mode == DH_SCALE_P2P) {
            const int tick = p.tick_base + step + 1;
            const int r = threadIdx.x;
            if (r < p.world) {
                volatile unsigned long long* dst = mb_scale_slot(p.peers[r], tick, p.rank);
                dst[0] = (unsigned long long)mine.hi;
                dst[1] = mine.lo;
                __threadfence_system();
                *mb_scale_flag(p.peers[r], tick, p.rank) = tick;
                mb_wait(mb_scale_flag(p.mailbox, tick, r), tick, p.halo_timeout_ms, p.status);
                const volatile unsigned long long* src = mb_scale_slot(p.mailbox, tick, r);
                s_parts[r].hi = (long long)src[0];
                s_parts[r].lo = src[1];
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            if (p.scale_mode == DH_SCALE_DEFERRED) {
                p.scale_part[0] = (unsigned long long)mine.hi;
                p.scale_part[1] = mine.lo;
            } else {
                Fx128 tot = mine;
                if (p.scale_mode == DH_SCALE_P2P) {
                    tot.hi = 0; tot.lo = 0ull;
                    for (int r = 0; r < p.world; r++) tot = fx_add(tot, s_parts[r]);
                }
                float step_sz, bc2s;
                adam_bias(step + 1, p.lr, &step_sz, &bc2s);
                adam_update(p.scale, &p.adam_mv_scale[0], &p.adam_mv_scale[1], (float)fx_to_double(tot), step_sz, bc2s);
            }
        }
    }
    if (threadIdx.x == 0) {
        if (p.iter_ns != nullptr && step < p.max_iters) p.iter_ns[2 * (size_t)step + 1] = globaltimer_ns();
        *p.step = step + 1;
    }
}

// DH_SCALE_DEFERRED: the partials of all ranks (gathered by the host side) -> Adam step of the scale.
__global__ void k_scale_apply(const dh_jointopt p, const unsigned long long* __restrict__ parts, int world) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Fx128 tot;
    tot.hi = 0; tot.lo = 0ull;
    for (int r = 0; r < world; r++) {
        Fx128 t;
        t.hi = (long long)parts[2 * r];
        t.lo = parts[2 * r + 1];
        tot = fx_add(tot, t);
    }
    float step_sz, bc2s;
    adam_bias(*p.step, p.lr, &step_sz, &bc2s);   // *step was already advanced by the iteration's k_finalize
    adam_update(p.scale, &p.adam_mv_scale[0], &p.adam_mv_scale[1], (float)fx_to_double(tot), step_sz, bc2s);
}

// ------------------------------------------------------------------------------------------------ host side
int check_sil(const dh_sil* s) {
    DH_REQUIRE(s != nullptr, "dh_sil is NULL");
    DH_REQUIRE(s->B > 0 && s->V > 0 && s->F > 0 && s->S > 0, "B, V, F, S must be positive");
    const int is = raster_size(*s);
    if (s->S % 32 != 0 || is > kMaxIS)
        return fail(DH_ERR_UNSUPPORTED, "S=%d aa=%d: S must be a multiple of 32 and S*(aa?2:1) <= %d", s->S, s->aa,
                    kMaxIS);
    DH_REQUIRE(s->faces && s->K && s->proj && s->bin_count && s->bins && s->fidx && s->alpha_bits && s->pos_pool &&
                   s->neg_pool && s->gmax && s->owned && s->negT && s->row_rng && s->neg_lists,
               "dh_sil has a NULL buffer");
    DH_REQUIRE(s->B <= 65535, "B > 65535 frames per call (grid.y limit); shard the sequence");
    return DH_OK;
}

size_t raster_smem_bytes(const dh_sil& s) {  // z-buffer strip + (small) face-owns-a-pixel bitmap + per-warp sets
    const int words = (2 * s.F + 31) / 32;
    return (size_t)kSH * raster_size(s) * sizeof(unsigned long long) +
           (size_t)(((words <= kOwnedSmemWords ? words : 0) + 3) / 4 * 4) * sizeof(uint32_t) +
           (size_t)kRasterWarps * sizeof(RasterWarp) + (size_t)raster_size(s) * sizeof(float) +
           (size_t)(s.aa ? kSH / 2 : kSH) * s.S;   // + the strip's target-mask cells
}

// pixels per frame the list path takes (dh_tune_set knob 0 lowers it: the tests force the bitmap path with it)
int g_neg_list_cap = kNLCap;
// (8 entries short of the block: the out scan's look-ahead read may touch the entry after the last one)
int neg_list_cap() { return g_neg_list_cap < kNLCap - 8 ? (g_neg_list_cap < 0 ? 0 : g_neg_list_cap) : kNLCap - 8; }

size_t bwd_smem_bytes(const dh_sil& s) {
    const int is = raster_size(s);
    return kBwdWarps * sizeof(BwdWarp) + (size_t)(2 * is * (is / 32) + s.S * ((s.S + 31) / 32)) * sizeof(uint32_t);
}
size_t bwd_lists_smem_bytes(const dh_sil& s) {
    const int is = raster_size(s);
    return kBwdWarps * (sizeof(BwdWarp) + sizeof(BwdSpans)) +
           (DH_ALPHA_GLOBAL ? 16 : (size_t)(is * (is / 32)) * sizeof(uint32_t)) +
           (DH_LISTS_GLOBAL ? 0 : (size_t)2 * kNLAxis * sizeof(uint16_t));
}
// CTAs per frame of k_neg_maps: 2 = one per side (rows / columns), 1 = both sides in one CTA (DH_NEG_SIDES overrides)
int neg_maps_sides() {
    const char* e = getenv("DH_NEG_SIDES");
    return (e != nullptr && e[0] == '2') ? 2 : ((e != nullptr && e[0] == '1') ? 1 : 2);
}
size_t neg_maps_smem_bytes(const dh_sil& s) {
    const int is = raster_size(s);
    return (size_t)(is * (is / 32 + 1) + is * (is / 32)) * sizeof(uint32_t);
}

template <typename KernelT>
int set_smem(KernelT kernel, size_t bytes) {
    // always: static + dynamic shared memory together may exceed the 48 KB default even when `bytes` does not
    DH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return DH_OK;
}

int launch_forward_common(const dh_sil& s, cudaStream_t st) {
    const int is = raster_size(s);
    dim3 gb((s.F + kThreads - 1) / kThreads, s.B);
    k_setup_bin<<<gb, kThreads, 0, st>>>(reinterpret_cast<const float4*>(s.proj), s.faces, s.V, s.F, is, is / kSH,
                                         s.bin_count, s.bins);
    DH_LAUNCH_OK("k_setup_bin");
    return DH_OK;
}

// Optional per-kernel timing: 9 events bracket the kernels of one iteration (dh_jointopt_profile).
struct IterEvents { cudaEvent_t ev[9]; };
// DH_CORR_FORK=0 in the environment keeps the correspondence kernel in line with the others (measurement knob)
// 0 in line, 1 branch after k_pose_prep, 2 branch after k_raster, 3 the same at low priority (DH_CORR_FORK overrides).
// Measured (300 frames, profiles/r2_jointopt_ncu.md): beside k_neg_maps a 10k-correspondence launch is free (-0.019 ms
// per iteration); a 50k one outlasts k_neg_maps, takes shared memory from the first wave of backward CTAs and costs
// +0.05 ms, so it stays in line.
int corr_fork_mode(int C) {
    const char* e = getenv("DH_CORR_FORK");
    if (e != nullptr && e[0] >= '0' && e[0] <= '3') return e[0] - '0';
    return C <= 20000 ? 2 : 0;
}
#define DH_REC(i) do { if (evs) cudaEventRecord(evs->ev[i], st); } while (0)

// The silhouette term's kernels of one iteration for the plan's frames: projection, binning, raster (+ fused loss
// epilogue) and, unless forward_only, the per-frame map kernel and the two backward kernels.
// part: 0 = everything, 1 = projection / binning / raster only, 2 = map kernel + backward only (pipelined groups).
int launch_sil_kernels(const dh_jointopt& p, bool forward_only, cudaStream_t st, IterEvents* evs = nullptr,
                       int part = 0) {
    const dh_sil& s = p.sil;
    const int is = raster_size(s), nstrips = is / kSH, B = s.B;
    dim3 gv((s.V + kThreads - 1) / kThreads, B);
    const bool stage1 = p.loss_mode == DH_LOSS_STAGE1;
    // dL/drend = gcoef * (k/2), gcoef = (lw / B) / keep_sum in fp32 like autograd (losses.py:69-75)
    const float gcoef = ((float)p.lw_sil / (float)p.B_total) / (float)p.keep_sum;
    int rc = DH_OK;
    if (part != 2) {
    k_project<true><<<gv, kThreads, 0, st>>>(p.verts_og, p.Rmat, p.trans, p.scale, s.K, s.orig_size,
                                              reinterpret_cast<float4*>(s.proj), s.V, s.bin_count, nstrips,
                                              p.loss_counts, s.owned, (2 * s.F + 31) / 32,
                                              stage1 ? p.offscreen : nullptr, (float)p.lw_offscreen, s.far_);
    DH_LAUNCH_OK("k_project");
    DH_REC(2);
    rc = launch_forward_common(s, st);
    if (rc) return rc;
    DH_REC(3);
    const size_t zb = raster_smem_bytes(s);
    rc = set_smem(k_raster<true>, zb);
    if (rc) return rc;
    k_raster<true><<<dim3(nstrips, B), kRasterThreads, zb, st>>>(s, p.mask_tri, gcoef, nullptr, p.loss_counts);
    DH_LAUNCH_OK("k_raster");
    DH_REC(4);
    }
    if (forward_only || part == 1) return DH_OK;
    rc = set_smem(k_neg_maps, neg_maps_smem_bytes(s));
    if (rc) return rc;
    float* fcoef = stage1 ? p.frame_coef : nullptr;
    k_neg_maps<<<dim3(B, neg_maps_sides()), kNegThreads, neg_maps_smem_bytes(s), st>>>(s, 1, neg_list_cap(), p.loss_counts, fcoef,
                                                                        (float)p.lw_sil);
    DH_LAUNCH_OK("k_neg_maps");
    const size_t sl = bwd_lists_smem_bytes(s), sb = bwd_smem_bytes(s);
    rc = set_smem(k_backward<true, true>, sl);
    if (rc) return rc;
    rc = set_smem(k_backward<true, false>, sb);
    if (rc) return rc;
    k_backward<true, true><<<dim3(p.nchunks, B), kBwdThreads, sl, st>>>(
        s, p.verts_og, p.Rmat, p.trans, p.scale, p.partials, nullptr, p.nchunks, gcoef, 0, fcoef, p.loss_counts);
    DH_LAUNCH_OK("k_backward<lists>");
    // frames with more contributing pixels than the lists hold (only these CTAs do any work)
    k_backward<true, false><<<dim3(p.nchunks, min(B, 32)), kBwdThreads, sb, st>>>(
        s, p.verts_og, p.Rmat, p.trans, p.scale, p.partials, nullptr, p.nchunks, gcoef, 1, fcoef, p.loss_counts);
    DH_LAUNCH_OK("k_backward<bitmaps>");
    return DH_OK;
}

// The plan restricted to frames [b0, b0 + nb) of its range: every per-frame array moved forward by b0 frames (the
// probe's blocks).  Only what launch_sil_kernels touches is meaningful in the result.
dh_jointopt sub_plan(const dh_jointopt& p, int b0, int nb) {
    dh_jointopt q = p;
    dh_sil& s = q.sil;
    const size_t o = (size_t)b0;
    const int is = raster_size(p.sil), nstrips = is / kSH, wpr = is >> 5, S = p.sil.S, wprp = (S + 31) >> 5;
    const int V = p.sil.V, F = p.sil.F;
    s.B = nb;
    s.K += o * 9;
    s.proj += o * V * 4;
    s.bin_count += o * nstrips * 2;
    s.bins += o * nstrips * 2 * F;
    s.fidx += o * is * is;
    s.alpha_bits += o * is * wpr;
    s.pos_pool += o * S * wprp;
    s.neg_pool += o * S * wprp;
    s.gmax += o;
    s.owned += o * ((2 * F + 31) / 32);
    s.negT += o * is * wpr;
    s.row_rng += o * 4 * is;
    s.neg_lists += o * 2 * kNLAxis;
    q.mask_tri += o * S * S;
    q.rot6d += o * 6;
    q.trans += o * 3;
    q.Rmat += o * 9;
    q.loss_counts += o * 4;
    q.partials += o * p.nchunks * 16;
    if (q.offscreen != nullptr) q.offscreen += o * 16;
    if (q.frame_coef != nullptr) q.frame_coef += o * 2;
    return q;
}

// A second stream + two events: the correspondence kernel (HBM-bound, reads only the poses) runs as its own branch
// beside the silhouette kernels (issue-bound) between k_pose_prep and k_pose_update.  Used when the iteration is
// captured into a graph; the plain-stream paths (profile, eval, grads, use_graph = 0) stay serial.
struct SideBranch { cudaStream_t stream; cudaEvent_t fork, join; bool after_raster; };

// half: 0 = the whole iteration; 1 = pose preparation + silhouette kernels only, 2 = correspondence kernel + pose
// update + bookkeeping only (dh_jointopt_run_part).
int launch_iteration(const dh_jointopt& p, int mode, float* g_rot, float* g_trans, float* g_scale,
                     cudaStream_t st, IterEvents* evs = nullptr, const SideBranch* side = nullptr, int half = 0) {
    const dh_sil& s = p.sil;
    const int B = s.B;
    const bool with_sil = p.lw_sil > 0.0;
    const bool with_corr = p.corr.records != nullptr && p.corr.lw_corr > 0.0;
    if (half != 0) {
        if (half == 1) {
            k_pose_prep<<<(B + 127) / 128, 128, 0, st>>>(p);
            DH_LAUNCH_OK("k_pose_prep");
            if (with_sil) return launch_sil_kernels(p, mode == 2, st);
            return DH_OK;
        }
        if (with_corr) {
            const int rc = launch_corr(p.corr.records, B, p.corr.C, p.Rmat, p.trans, p.scale, s.K, s.S, p.corr.delta,
                                       p.corr.partials, p.corr.nslots, st);
            if (rc) return rc;
        }
        k_pose_update<<<(B + kPoseUpdateThreads / 32 - 1) / (kPoseUpdateThreads / 32), kPoseUpdateThreads, 0, st>>>(p, mode, g_rot, g_trans, with_sil ? 1 : 0, with_corr ? 1 : 0);
        DH_LAUNCH_OK("k_pose_update");
        k_finalize<<<1, kThreads, 0, st>>>(p, mode, g_scale);
        DH_LAUNCH_OK("k_finalize");
        return DH_OK;
    }
    DH_REC(0);
    k_pose_prep<<<(B + 127) / 128, 128, 0, st>>>(p);
    DH_LAUNCH_OK("k_pose_prep");
    DH_REC(1);
    const bool forked = with_corr && with_sil && side != nullptr;
    auto corr_launch = [&](bool fork) -> int {
        cudaStream_t cs = st;
        if (fork) {
            DH_CUDA(cudaEventRecord(side->fork, st));
            DH_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
            cs = side->stream;
        }
        const int rc = launch_corr(p.corr.records, B, p.corr.C, p.Rmat, p.trans, p.scale, s.K, s.S, p.corr.delta,
                                   p.corr.partials, p.corr.nslots, cs);
        if (rc) return rc;
        if (fork) DH_CUDA(cudaEventRecord(side->join, cs));
        return DH_OK;
    };
    if (with_corr && !(forked && side->after_raster)) {
        const int rc = corr_launch(forked);
        if (rc) return rc;
    }
    DH_REC(8);
    if (with_sil) {
        int rc;
        if (forked && side->after_raster) {   // the branch starts beside k_neg_maps, which leaves most of the machine idle
            rc = launch_sil_kernels(p, mode == 2, st, evs, 1);
            if (!rc) rc = corr_launch(true);
            if (!rc) rc = launch_sil_kernels(p, mode == 2, st, evs, 2);
        } else {
            rc = launch_sil_kernels(p, mode == 2, st, evs);
        }
        if (rc) return rc;
    } else {
        DH_REC(2); DH_REC(3); DH_REC(4);
    }
    DH_REC(5);
    if (forked) DH_CUDA(cudaStreamWaitEvent(st, side->join, 0));
    k_pose_update<<<(B + kPoseUpdateThreads / 32 - 1) / (kPoseUpdateThreads / 32), kPoseUpdateThreads, 0, st>>>(p, mode, g_rot, g_trans, with_sil ? 1 : 0, with_corr ? 1 : 0);
    DH_LAUNCH_OK("k_pose_update");
    DH_REC(6);
    k_finalize<<<1, kThreads, 0, st>>>(p, mode, g_scale);
    DH_LAUNCH_OK("k_finalize");
    DH_REC(7);
    return DH_OK;
}

int check_plan(const dh_jointopt* p) {
    DH_REQUIRE(p != nullptr, "dh_jointopt is NULL");
    int rc = check_sil(&p->sil);
    if (rc) return rc;
    DH_REQUIRE(p->sil.gpool != nullptr, "sil.gpool is NULL");
    DH_REQUIRE(p->verts_og && p->mask_tri && p->rot6d && p->trans && p->scale && p->step && p->hist && p->moments,
               "dh_jointopt has a NULL parameter/state pointer");
    DH_REQUIRE(p->adam_m_rot && p->adam_v_rot && p->adam_m_trans && p->adam_v_trans && p->adam_mv_scale,
               "dh_jointopt has a NULL Adam state pointer");
    DH_REQUIRE(p->Rmat && p->smooth_terms && p->loss_counts && p->partials && p->frame_terms,
               "dh_jointopt has a NULL scratch pointer");
    DH_REQUIRE(p->nchunks >= 1 && (p->sil.F + p->nchunks - 1) / p->nchunks <= kChunkFaces,
               "nchunks must be >= ceil(F / 1024)");
    DH_REQUIRE(p->B_total >= p->sil.B, "B_total < B");
    DH_REQUIRE(p->max_iters >= 0, "max_iters < 0");
    DH_REQUIRE(p->loss_mode == DH_LOSS_JOINT || p->loss_mode == DH_LOSS_STAGE1, "bad loss_mode");
    if (p->loss_mode == DH_LOSS_STAGE1) {
        DH_REQUIRE(p->sil.aa == 0, "stage-1 IoU loss: the reference renders it without anti-aliasing");
        DH_REQUIRE(p->offscreen != nullptr && p->frame_coef != nullptr, "stage-1: offscreen / frame_coef are NULL");
        DH_REQUIRE(!(p->lw_smooth > 0.0) && !p->optimize_scale, "stage-1: no smoothness term, no scale");
    }
    DH_REQUIRE(p->scale_mode == DH_SCALE_LOCAL || p->scale_mode == DH_SCALE_P2P || p->scale_mode == DH_SCALE_DEFERRED,
               "bad scale_mode");
    if (p->optimize_scale && p->scale_mode == DH_SCALE_P2P) {
        DH_REQUIRE(p->mailbox != nullptr && p->world >= 1 && p->world <= DH_MAX_RANKS && p->rank >= 0 &&
                       p->rank < p->world, "scale_mode P2P needs a mailbox and 1 <= world <= DH_MAX_RANKS");
        for (int r = 0; r < p->world; r++) DH_REQUIRE(p->peers[r] != nullptr, "scale_mode P2P: peers[%d] is NULL", r);
    }
    if (p->optimize_scale && p->scale_mode == DH_SCALE_DEFERRED)
        DH_REQUIRE(p->scale_part != nullptr, "scale_mode DEFERRED needs scale_part");
    DH_REQUIRE(p->keep_sum > 0.0 || !(p->lw_sil > 0.0), "keep_sum must be positive");
    if (p->corr.records != nullptr && p->corr.lw_corr > 0.0) {
        DH_REQUIRE(p->corr.partials != nullptr && p->corr.C > 0 && p->corr.nslots > 0, "corr: bad plan");
        DH_REQUIRE((p->corr.w_sum_dev != nullptr || p->corr.w_sum > 0.0) && p->corr.delta > 0.0f,
                   "corr: w_sum and delta must be positive");
    }
    return DH_OK;
}

// Cached graph executables, keyed by the whole plan (every pointer and scalar a captured launch bakes in), the
// device it was captured on and the tuning knobs read at capture time.  Guarded by a mutex (one host thread per GPU
// is the convention, but several GPUs may be driven from one process); bounded: the oldest entry goes first.
struct GraphEntry {
    dh_jointopt plan;
    int device;
    int list_cap;
    int fork_mode;
    cudaGraphExec_t exec;
};
constexpr size_t kGraphCacheMax = 32;
std::mutex& graph_mutex() {
    static std::mutex m;
    return m;
}
std::vector<GraphEntry>& graph_cache() {
    static std::vector<GraphEntry> c;
    return c;
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

int dh_sil_scratch_bytes(int32_t B, int32_t V, int32_t F, int32_t S, int32_t aa, int64_t* out13) {
    DH_REQUIRE(out13 != nullptr && B > 0 && V > 0 && F > 0 && S > 0, "bad arguments");
    out13[8] = (int64_t)B * 4;
    out13[9] = (int64_t)B * ((2 * (int64_t)F + 31) / 32) * 4;
    const int64_t is = aa ? 2 * S : S;
    const int64_t nstrips = (is + kSH - 1) / kSH, wprp = (S + 31) / 32;
    out13[0] = (int64_t)B * V * 4 * 4;
    out13[1] = (int64_t)B * nstrips * 2 * 4;
    out13[2] = (int64_t)B * nstrips * 2 * F * 4;
    out13[3] = (int64_t)B * is * is * 4;
    out13[4] = (int64_t)B * is * (is / 32) * 4;
    out13[5] = (int64_t)B * S * wprp * 4;
    out13[6] = out13[5];
    out13[7] = 16;   // gpool: the fused path no longer materialises dL/drend, the API path reads the caller's grad_rend
    out13[10] = (int64_t)B * is * (is / 32) * 4;
    out13[11] = (int64_t)B * 4 * is * 2;
    out13[12] = (int64_t)B * 2 * kNLAxis * 2;
    return DH_OK;
}

int dh_sil_forward(const dh_sil* s, const float* verts_cam, float* rend, void* stream) {
    int rc = check_sil(s);
    if (rc) return rc;
    DH_REQUIRE(verts_cam && rend, "NULL verts_cam / rend");
    cudaStream_t st = (cudaStream_t)stream;
    const int is = raster_size(*s), nstrips = is / kSH;
    dim3 gv((s->V + kThreads - 1) / kThreads, s->B);
    k_project<false><<<gv, kThreads, 0, st>>>(verts_cam, nullptr, nullptr, nullptr, s->K, s->orig_size,
                                               reinterpret_cast<float4*>(s->proj), s->V, s->bin_count, nstrips,
                                               nullptr, s->owned, (2 * s->F + 31) / 32);
    DH_LAUNCH_OK("k_project");
    rc = launch_forward_common(*s, st);
    if (rc) return rc;
    const size_t zb = raster_smem_bytes(*s);
    rc = set_smem(k_raster<false>, zb);
    if (rc) return rc;
    k_raster<false><<<dim3(nstrips, s->B), kRasterThreads, zb, st>>>(*s, nullptr, 0.0f, rend, nullptr);
    DH_LAUNCH_OK("k_raster");
    return DH_OK;
}

int dh_sil_backward(const dh_sil* s, const float* verts_cam, const float* grad_rend, float* grad_verts,
                    void* stream) {
    int rc = check_sil(s);
    if (rc) return rc;
    DH_REQUIRE(verts_cam && grad_rend && grad_verts, "NULL verts_cam / grad_rend / grad_verts");
    cudaStream_t st = (cudaStream_t)stream;
    const long long ncell = (long long)s->B * s->S * s->S;
    DH_CUDA(cudaMemsetAsync(grad_verts, 0, (size_t)s->B * s->V * 3 * sizeof(float), st));
    DH_CUDA(cudaMemsetAsync(s->gmax, 0, (size_t)s->B * sizeof(float), st));
    k_grad_signs<<<(unsigned)((ncell + kThreads - 1) / kThreads), kThreads, 0, st>>>(
        grad_rend, s->pos_pool, s->neg_pool, s->gmax, ncell, s->S * s->S, s->aa ? 0.25f : 1.0f);
    DH_LAUNCH_OK("k_grad_signs");
    dh_sil t = *s;
    t.gpool = const_cast<float*>(grad_rend);
    rc = set_smem(k_neg_maps, neg_maps_smem_bytes(t));
    if (rc) return rc;
    k_neg_maps<<<dim3(t.B, 2), kNegThreads, neg_maps_smem_bytes(t), st>>>(t, 0, 0, nullptr, nullptr, 0.0f);
    DH_LAUNCH_OK("k_neg_maps");
    const size_t sb = bwd_smem_bytes(t);
    rc = set_smem(k_backward<false, false>, sb);
    if (rc) return rc;
    const int nchunks = dh_jointopt_default_chunks(s->B, s->F);
    k_backward<false, false><<<dim3(nchunks, s->B), kBwdThreads, sb, st>>>(
        t, verts_cam, nullptr, nullptr, nullptr, nullptr, grad_verts, nchunks, 0.0f, 0, nullptr, nullptr);
    DH_LAUNCH_OK("k_backward");
    return DH_OK;
}

int dh_bwd_schedule(int32_t n_items, int32_t* starts_out, int32_t cap) {
    DH_REQUIRE(n_items >= 0 && n_items <= 2 * kChunkFaces && starts_out != nullptr && cap >= kMaxBatches + 1,
               "dh_bwd_schedule: 0 <= n_items <= %d, cap >= %d", 2 * kChunkFaces, kMaxBatches + 1);
    uint16_t starts[kMaxBatches + 1];
#if DH_GUIDED
    const int nb = bwd_guided_schedule(n_items, kBwdWarps, DH_GUIDE_MIN, DH_GUIDE_DIV, kMaxBatches, starts);
#else
    int nb = 0;
    for (int pos = 0; pos < n_items; pos += 32) starts[nb++] = (uint16_t)pos;
    starts[nb] = (uint16_t)n_items;
#endif
    for (int i = 0; i <= nb; i++) starts_out[i] = starts[i];
    return nb;
}

int dh_tune_set(int32_t knob, int32_t value) {
    DH_REQUIRE(knob == 0, "dh_tune_set: unknown knob");
    g_neg_list_cap = value < 0 ? kNLCap : value;  // < 0 restores the default
    return DH_OK;   // cached graphs carry the value they were captured with in their key: no stale reuse
}

int dh_rot6d_to_matrix(const float* rot6d, float* R, int32_t B, void* stream) {
    DH_REQUIRE(rot6d && R && B > 0, "bad arguments");
    k_rot6d_to_matrix<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rot6d, R, B);
    DH_LAUNCH_OK("k_rot6d_to_matrix");
    return DH_OK;
}

int dh_transform_verts(const float* verts, const float* R, const float* T, const float* scale, float* out,
                       int32_t B, int32_t V, void* stream) {
    DH_REQUIRE(verts && R && T && scale && out && B > 0 && V > 0 && B <= 65535, "bad arguments");
    k_transform_verts<<<dim3((V + kThreads - 1) / kThreads, B), kThreads, 0, (cudaStream_t)stream>>>(verts, R, T,
                                                                                                     scale, out, V);
    DH_LAUNCH_OK("k_transform_verts");
    return DH_OK;
}

int dh_masks_prepare(const float* target_masks, int8_t* tri, unsigned long long* keep_count, int64_t n,
                     void* stream) {
    DH_REQUIRE(target_masks && tri && keep_count && n > 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    DH_CUDA(cudaMemsetAsync(keep_count, 0, sizeof(unsigned long long), st));
    long long blocks = (n + kThreads - 1) / kThreads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_masks_prepare<<<(unsigned)blocks, kThreads, 0, st>>>(target_masks, tri, keep_count, (long long)n);
    DH_LAUNCH_OK("k_masks_prepare");
    return DH_OK;
}

int dh_mesh_moments(const float* verts, int32_t V, double* out12, void* stream) {
    DH_REQUIRE(verts && out12 && V > 0, "bad arguments");
    k_mesh_moments<<<1, kThreads, 0, (cudaStream_t)stream>>>(verts, V, out12);
    DH_LAUNCH_OK("k_mesh_moments");
    return DH_OK;
}

int dh_adam_step(float* param, const float* grad, float* m, float* v, int64_t n, double lr, int32_t t,
                 void* stream) {
    DH_REQUIRE(param && grad && m && v && n > 0 && t >= 1, "bad arguments");
    float step_size, bc2s;
    adam_bias(t, lr, &step_size, &bc2s);
    k_adam<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(param, grad, m, v,
                                                                                           (long long)n, step_size,
                                                                                           bc2s);
    DH_LAUNCH_OK("k_adam");
    return DH_OK;
}

int dh_jointopt_scratch_bytes(int32_t B, int32_t nchunks, int64_t* out5) {
    DH_REQUIRE(out5 != nullptr && B > 0 && nchunks > 0, "bad arguments");
    out5[0] = (int64_t)B * 9 * 4;
    out5[1] = (int64_t)B * 16 * 8;
    out5[2] = (int64_t)B * 4 * 4;
    out5[3] = (int64_t)B * nchunks * 16 * 4;
    out5[4] = (int64_t)B * 8 * 8;
    return DH_OK;
}

int dh_jointopt_default_chunks(int32_t B, int32_t F) {
    (void)B;
    int c = (F + 4 * kThreads - 1) / (4 * kThreads);  // ~4 faces per thread
    const int cmin = (F + kChunkFaces - 1) / kChunkFaces;  // the CTA's item list holds kChunkFaces faces
    if (c < cmin) c = cmin;
    if (c < 1) c = 1;
    return c;
}

int dh_jointopt_run(const dh_jointopt* p, int32_t n_iters, int32_t use_graph, void* stream) {
    int rc = check_plan(p);
    if (rc) return rc;
    DH_REQUIRE(n_iters >= 0, "n_iters < 0");
    cudaStream_t st = (cudaStream_t)stream;
    if (!use_graph) {
        for (int it = 0; it < n_iters; it++) {
            rc = launch_iteration(*p, 0, nullptr, nullptr, nullptr, st);
            if (rc) return rc;
        }
        return DH_OK;
    }
    cudaGraphExec_t exec = nullptr;
    int device = 0;
    DH_CUDA(cudaGetDevice(&device));
    const int list_cap = neg_list_cap();
    const int fork_mode = corr_fork_mode(p->corr.C);
    std::lock_guard<std::mutex> lock(graph_mutex());
    for (auto& e : graph_cache())
        if (e.device == device && e.list_cap == list_cap && e.fork_mode == fork_mode &&
            memcmp(&e.plan, p, sizeof(dh_jointopt)) == 0)
            exec = e.exec;
    if (exec == nullptr) {
        // warm the function attributes outside capture
        rc = set_smem(k_raster<true>, raster_smem_bytes(p->sil));
        if (rc) return rc;
        rc = set_smem(k_backward<true, false>, bwd_smem_bytes(p->sil));
        if (rc) return rc;
        rc = set_smem(k_backward<true, true>, bwd_lists_smem_bytes(p->sil));
        if (rc) return rc;
        rc = set_smem(k_neg_maps, neg_maps_smem_bytes(p->sil));
        if (rc) return rc;
        cudaStream_t cs;
        DH_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        SideBranch side;
        const bool fork_corr = fork_mode != 0;
        side.after_raster = fork_mode >= 2;
        if (fork_corr) {
            int pr_lo = 0, pr_hi = 0;
            DH_CUDA(cudaDeviceGetStreamPriorityRange(&pr_lo, &pr_hi));
            DH_CUDA(cudaStreamCreateWithPriority(&side.stream, cudaStreamNonBlocking, fork_mode == 3 ? pr_lo : 0));
            DH_CUDA(cudaEventCreateWithFlags(&side.fork, cudaEventDisableTiming));
            DH_CUDA(cudaEventCreateWithFlags(&side.join, cudaEventDisableTiming));
        }
        cudaGraph_t graph = nullptr;
        DH_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        rc = launch_iteration(*p, 0, nullptr, nullptr, nullptr, cs, nullptr, fork_corr ? &side : nullptr);
        cudaError_t ce = cudaStreamEndCapture(cs, &graph);
        if (fork_corr) {
            cudaEventDestroy(side.fork);
            cudaEventDestroy(side.join);
            cudaStreamDestroy(side.stream);
        }
        if (rc || ce != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            cudaStreamDestroy(cs);
            return rc ? rc : fail(DH_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
        }
        ce = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        cudaStreamDestroy(cs);
        if (ce != cudaSuccess) return fail(DH_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ce));
        auto& c = graph_cache();
        if (c.size() >= kGraphCacheMax) {
            cudaGraphExecDestroy(c.front().exec);
            c.erase(c.begin());
        }
        GraphEntry e;
        memcpy(&e.plan, p, sizeof(dh_jointopt));
        e.device = device;
        e.list_cap = list_cap;
        e.fork_mode = fork_mode;
        e.exec = exec;
        c.push_back(e);
    }
    for (int it = 0; it < n_iters; it++) DH_CUDA(cudaGraphLaunch(exec, st));
    return DH_OK;
}

int dh_jointopt_run_part(const dh_jointopt* p, int32_t part, void* stream) {
    int rc = check_plan(p);
    if (rc) return rc;
    DH_REQUIRE(part == 1 || part == 2, "part must be 1 or 2");
    if (part == 1) {   // function attributes (the graph path sets them before its capture)
        rc = set_smem(k_raster<true>, raster_smem_bytes(p->sil));
        if (rc) return rc;
    }
    return launch_iteration(*p, 0, nullptr, nullptr, nullptr, (cudaStream_t)stream, nullptr, nullptr, part);
}

int dh_jointopt_eval(const dh_jointopt* p, void* stream) {
    int rc = check_plan(p);
    if (rc) return rc;
    return launch_iteration(*p, 2, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int dh_jointopt_grads(const dh_jointopt* p, float* grad_rot6d, float* grad_trans, float* grad_scale, void* stream) {
    int rc = check_plan(p);
    if (rc) return rc;
    DH_REQUIRE(grad_rot6d && grad_trans, "NULL gradient outputs");
    return launch_iteration(*p, 1, grad_rot6d, grad_trans, grad_scale, (cudaStream_t)stream);
}

int dh_jointopt_profile(const dh_jointopt* p, int32_t n_iters, float* ms_out_host, void* stream) {
    int rc = check_plan(p);
    if (rc) return rc;
    DH_REQUIRE(n_iters > 0 && ms_out_host != nullptr, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    IterEvents evs;
    for (int i = 0; i < 9; i++) DH_CUDA(cudaEventCreate(&evs.ev[i]));
    for (int i = 0; i < 8; i++) ms_out_host[i] = 0.0f;
    // event order on the stream: 0 pose_prep 1 corr 8 project 2 setup_bin 3 raster 4 backward 5 pose_update 6 finalize 7
    const int seg[8][2] = {{0, 1}, {8, 2}, {2, 3}, {3, 4}, {4, 5}, {5, 6}, {6, 7}, {1, 8}};
    for (int it = 0; it < n_iters; it++) {
        rc = launch_iteration(*p, 0, nullptr, nullptr, nullptr, st, &evs);
        if (rc) break;
        cudaError_t e = cudaEventSynchronize(evs.ev[7]);
        if (e != cudaSuccess) { rc = fail(DH_ERR_CUDA, "profile sync: %s", cudaGetErrorString(e)); break; }
        for (int i = 0; i < 8; i++) {
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, evs.ev[seg[i][0]], evs.ev[seg[i][1]]);
            ms_out_host[i] += ms / (float)n_iters;
        }
    }
    for (int i = 0; i < 9; i++) cudaEventDestroy(evs.ev[i]);
    return rc;
}

int dh_jointopt_probe(const dh_jointopt* p, int32_t nblocks, float* ms_out_host, void* stream) {
    int rc = check_plan(p);
    if (rc) return rc;
    DH_REQUIRE(nblocks >= 1 && nblocks <= p->sil.B && nblocks <= 256 && ms_out_host != nullptr, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    dh_jointopt q = *p;
    q.mailbox = nullptr;   // the probe never waits on a neighbour: halo_prev / halo_next as given (may be NULL)
    q.peer_prev = q.peer_next = nullptr;
    const int B = q.sil.B;
    const bool with_sil = q.lw_sil > 0.0;
    const bool with_corr = q.corr.records != nullptr && q.corr.lw_corr > 0.0;
    std::vector<cudaEvent_t> ev(nblocks + 3);
    for (auto& e : ev) DH_CUDA(cudaEventCreate(&e));
    k_pose_prep<<<(B + 127) / 128, 128, 0, st>>>(q);
    DH_LAUNCH_OK("k_pose_prep");
    // one untimed pass over everything (first-touch, function attributes), then block by block
    if (with_sil) rc = launch_sil_kernels(q, false, st);
    for (int pass = 0; pass < 2 && !rc && with_corr; pass++) {
        if (pass == 1) cudaEventRecord(ev[nblocks + 1], st);
        rc = launch_corr(q.corr.records, B, q.corr.C, q.Rmat, q.trans, q.scale, q.sil.K, q.sil.S, q.corr.delta,
                         q.corr.partials, q.corr.nslots, st);
        if (pass == 1) cudaEventRecord(ev[nblocks + 2], st);
    }
    for (int k = 0; k < nblocks && !rc; k++) {
        const int b0 = (int)((long long)B * k / nblocks), b1 = (int)((long long)B * (k + 1) / nblocks);
        cudaEventRecord(ev[k], st);
        if (with_sil) rc = launch_sil_kernels(sub_plan(q, b0, b1 - b0), false, st);
    }
    cudaEventRecord(ev[nblocks], st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (!rc && e != cudaSuccess) rc = fail(DH_ERR_CUDA, "probe sync: %s", cudaGetErrorString(e));
    for (int k = 0; k <= nblocks; k++) ms_out_host[k] = 0.0f;
    if (!rc) {
        for (int k = 0; k < nblocks; k++) cudaEventElapsedTime(&ms_out_host[k], ev[k], ev[k + 1]);
        if (with_corr) cudaEventElapsedTime(&ms_out_host[nblocks], ev[nblocks + 1], ev[nblocks + 2]);
    }
    for (auto& x : ev) cudaEventDestroy(x);
    return rc;
}

int dh_scale_apply(const dh_jointopt* p, const unsigned long long* parts_dev, int32_t world, void* stream) {
    int rc = check_plan(p);
    if (rc) return rc;
    DH_REQUIRE(parts_dev != nullptr && world >= 1, "bad arguments");
    k_scale_apply<<<1, 32, 0, (cudaStream_t)stream>>>(*p, parts_dev, world);
    DH_LAUNCH_OK("k_scale_apply");
    return DH_OK;
}

int dh_jointopt_release(const dh_jointopt* p) {
    std::lock_guard<std::mutex> lock(graph_mutex());
    auto& c = graph_cache();
    for (size_t i = 0; i < c.size();) {
        if (p == nullptr || memcmp(&c[i].plan, p, sizeof(dh_jointopt)) == 0) {
            cudaGraphExecDestroy(c[i].exec);
            c.erase(c.begin() + i);
        } else {
            i++;
        }
    }
    return DH_OK;
}

}  // extern "C"
