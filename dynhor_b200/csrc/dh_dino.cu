// dh_dino.cu -- DINO template matching (pose_initializtion.py:295-311) on the sm_100a tensor cores.
//
// score[n, f] = sum_p m_f[p] <g_f[p], r_n[p]> / (|g_f[p]| |r_n[p]|) / sum_p m_f[p]   (pose_initializtion.py:295-296)
// is one dense contraction over K = P*D once both banks are pre-scaled (k_dino_prescale):
//     templ [N, K]  = r[n,p,:] / |r[n,p,:]|            frames [Fm, K] = m_f[p] g[f,p,:] / |g[f,p,:]| / sum_p m_f[p]
// k_dino_gemm   C[n, f] = templ . frames^T, bf16 x bf16 -> fp32:  TMA (128B swizzle) -> shared memory ring ->
//               tcgen05.mma (one elected thread, accumulators in TMEM) -> tcgen05.ld epilogue.  The output is only
//               N x Fm (1000 x 300) while K is ~5e5, so the grid splits K: one CTA per (128-template tile, 320-frame
//               tile, K slice), each streaming its slice once; partial tiles go to a small fp32 workspace.
// k_dino_reduce split-K reduction of the partial tiles into scores [Fm, N] (coalesced both ways via a smem transpose).
// k_dino_topk   per frame: top-k selection (torch.topk / argmax semantics of pose_initializtion.py:299,309;
//               ties -> lowest index).
// (Round 2 measured the two tail kernels merged into one -- a CTA per 8 frames reducing its 32-byte sectors of the
//  workspace and selecting from shared memory: 0.372 ms per batch against 0.312 ms, too few CTAs to hide the strided
//  reads -- and kept them apart.)
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

#include "dh_common.h"

namespace {

constexpr int BM = 128;        // templates per CTA tile (UMMA M, cta_group::1)
constexpr int BN = 160;        // frames per MMA (UMMA N), two such tiles per CTA
constexpr int NT = 2;
constexpr int BK = 64;         // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;           // 16 KB
constexpr int B_BYTES = BN * BK * 2;           // 20 KB
constexpr int STAGE_BYTES = A_BYTES + NT * B_BYTES;
constexpr int TMEM_COLS = 512;                 // power of two >= NT * BN
constexpr int GEMM_THREADS = 192;              // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment*/ + 256 /*barriers*/;
constexpr int MAX_TOPK = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a protocol error traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// K-major operand tile with 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO); sm_100 descriptor
// version 1, layout type 2 (SWIZZLE_128B).  cute/arch/mma_sm100_desc.hpp::SmemDescriptor.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset
    d |= (uint64_t)1 << 46;                 // descriptor version
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24.
__device__ __forceinline__ uint32_t umma_idesc() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// the same arrive delivered to the barrier at this offset in every CTA of `mask` (cluster multicast)
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%4, %5}], [%2], %3;"
        ::"r"(dst), "l"(map), "r"(bar), "h"(mask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// CS = CTAs per cluster along the template (M) tiles.  CS > 1: the CTAs of a cluster work on the same K slice and
// the same frame tile, so each loads 1/CS of the frame tile and multicasts it to the others -- the frame tile
// crosses L2 -> SM once per cluster instead of once per CTA (map_b's box is then 320/CS rows).
template <int CS>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_dino_gemm(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
            float* __restrict__ partial, int kblocks_total, int kblocks_per_slice, int m_pad, int ldc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), tfull = smem_u32(bars + 2 * STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * (NT * BN), slice = blockIdx.z;
    const int kb0 = slice * kblocks_per_slice;
    const int nkb = max(0, min(kblocks_total, kb0 + kblocks_per_slice) - kb0);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int i = 0; i < STAGES; i++) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, CS);   // every CTA of the cluster must have drained the stage
        }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CS > 1) cluster_sync_all();          // peers' barriers are initialised before anything is multicast
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    uint32_t cta_rank = 0;
    if (CS > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    const uint16_t mc_mask = (uint16_t)((1u << CS) - 1u);

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (one elected lane)
        if (lane == 0) {
            for (int kb = 0; kb < nkb; kb++) {
                const int st = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                mbar_wait(empty0 + 8 * st, ph ^ 1u);
                const uint32_t dst = smem_u32(smem + st * STAGE_BYTES);
                mbar_expect_tx(full0 + 8 * st, STAGE_BYTES);
                const int kc = (kb0 + kb) * BK;
                tma_load_2d(dst, &map_a, full0 + 8 * st, kc, m0);
                if (CS == 1) {
                    tma_load_2d(dst + A_BYTES, &map_b, full0 + 8 * st, kc, n0);
                    tma_load_2d(dst + A_BYTES + B_BYTES, &map_b, full0 + 8 * st, kc, n0 + BN);
                } else {
                    constexpr int RP = NT * BN / CS;   // rows of the 320-row frame tile this CTA fetches
                    tma_load_2d_mc(dst + A_BYTES + cta_rank * (RP * BK * 2), &map_b, full0 + 8 * st, kc,
                                   n0 + (int)cta_rank * RP, mc_mask);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (one elected lane)
        if (lane == 0) {
            const uint32_t idesc = umma_idesc();
            for (int kb = 0; kb < nkb; kb++) {
                const int st = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                mbar_wait(full0 + 8 * st, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(smem + st * STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; k++) {
                    const uint64_t da = umma_desc(a_addr + k * UMMA_K * 2);
#pragma unroll
                    for (int t = 0; t < NT; t++) {
                        const uint64_t db = umma_desc(a_addr + A_BYTES + t * B_BYTES + k * UMMA_K * 2);
                        umma_bf16(tmem_base + t * BN, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                }
                // frees the stage once these MMAs have read it (in every CTA of the cluster when multicasting)
                if (CS == 1) umma_commit(empty0 + 8 * st);
                else         umma_commit_mc(empty0 + 8 * st, mc_mask);
            }
            umma_commit(tfull);                 // accumulators complete
        }
    } else {
        // ------------------------------------------------------------------ epilogue: TMEM -> registers -> workspace
        mbar_wait(tfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                 // TMEM lane quadrant this warp may access
        const int row = m0 + q * 32 + lane;
        float* out = partial + ((size_t)slice * m_pad + row) * ldc + n0;
        for (int c = 0; c < NT * BN; c += 16) {
            uint32_t v[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c;
            if (nkb > 0) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                      "=r"(v[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = 0u;
            }
            float4* o4 = reinterpret_cast<float4*>(out + c);
#pragma unroll
            for (int i = 0; i < 4; i++)
                o4[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                    __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CS > 1) cluster_sync_all();          // nobody leaves while peers may still multicast into its smem
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// Split-K reduction: scores[f, n] = sum_slices partial[slice, n, f], through a 32x32 shared-memory transpose so that
// both the partial reads (frames contiguous) and the score writes (templates contiguous) are coalesced.
__global__ void __launch_bounds__(256)
k_dino_reduce(const float* __restrict__ partial, int nslices, int m_pad, int ldc, int N, int Fm,
              float* __restrict__ scores) {
    __shared__ float tile[32][33];
    const int n0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int nl = ty; nl < 32; nl += 8) {
        float acc = 0.0f;
        const int n = n0 + nl, f = f0 + tx;
        if (n < N && f < Fm)
            for (int s = 0; s < nslices; s++) acc += partial[((size_t)s * m_pad + n) * ldc + f];
        tile[nl][tx] = acc;
    }
    __syncthreads();
    for (int fl = ty; fl < 32; fl += 8) {
        const int f = f0 + fl, n = n0 + tx;
        if (f < Fm && n < N) scores[(size_t)f * N + n] = tile[tx][fl];
    }
}

// One CTA per frame: k rounds of block-wide arg-max over scores[f, :] (largest value, lowest index on ties), which
// is torch.topk(largest=True)'s order on tie-free data (pose_initializtion.py:299,309).
__global__ void __launch_bounds__(256)
k_dino_topk(const float* __restrict__ scores, int N, int k, float* __restrict__ topk_vals,
            int32_t* __restrict__ topk_idx) {
    extern __shared__ float s_scores[];  // [N]
    __shared__ float s_bv[8];
    __shared__ int s_bi[8];
    const int f = blockIdx.x, tid = threadIdx.x;
    for (int n = tid; n < N; n += 256) s_scores[n] = scores[(size_t)f * N + n];
    __syncthreads();
    for (int j = 0; j < k; j++) {
        float bv = -3.0e38f;
        int bi = 0x7FFFFFFF;
        for (int n = tid; n < N; n += 256) {
            const float v = s_scores[n];
            if (v > bv || (v == bv && n < bi)) { bv = v; bi = n; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if ((tid & 31) == 0) { s_bv[tid >> 5] = bv; s_bi[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 8; w++)
                if (s_bv[w] > bv || (s_bv[w] == bv && s_bi[w] < bi)) { bv = s_bv[w]; bi = s_bi[w]; }
            topk_vals[(size_t)f * k + j] = bv;
            topk_idx[(size_t)f * k + j] = bi;
            if (bi < N) s_scores[bi] = -3.0e38f;
        }
        __syncthreads();
    }
}

// feats [n, P, D] fp32 -> bf16 [n, P*D]: each patch vector divided by its L2 norm, times mask[n,p] / sum_p mask[n,:]
// when a mask is given.  One warp per (n, p) patch.
__global__ void __launch_bounds__(256)
k_dino_prescale(const float* __restrict__ feats, const float* __restrict__ mask, int n_rows, int P, int D,
                __nv_bfloat16* __restrict__ out) {
    const long long patch = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (patch >= (long long)n_rows * P) return;
    const int row = (int)(patch / P);
    const float* x = feats + patch * D;
    float ss = 0.0f;
    for (int d = lane; d < D; d += 32) ss += x[d] * x[d];
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    float w = 1.0f;
    if (mask != nullptr) {
        float msum = 0.0f;
        for (int p = lane; p < P; p += 32) msum += mask[(size_t)row * P + p];
        for (int o = 16; o > 0; o >>= 1) msum += __shfl_xor_sync(0xffffffffu, msum, o);
        w = mask[patch] / msum;
    }
    const float sc = w / fmaxf(sqrtf(ss), 1e-12f);
    __nv_bfloat16* o = out + patch * D;
    for (int d = lane; d < D; d += 32) o[d] = __float2bfloat16_rn(x[d] * sc);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t kdim, int box_rows) {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || p == nullptr || q != cudaDriverEntryPointSuccess)
            return dh::fail(DH_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
        fn = (EncodeTiledFn)p;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)kdim, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)kdim * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return dh::fail(DH_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return DH_OK;
}

template <int CS>
int launch_gemm_cluster(dim3 grid, cudaStream_t st, const CUtensorMap& map_a, const CUtensorMap& map_b, float* ws,
                        int kblocks, int kb_per_slice, int m_pad, int ldc) {
    DH_CUDA(cudaFuncSetAttribute(k_dino_gemm<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    if (CS == 8)
        DH_CUDA(cudaFuncSetAttribute(k_dino_gemm<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DH_CUDA(cudaLaunchKernelEx(&cfg, k_dino_gemm<CS>, map_a, map_b, ws, kblocks, kb_per_slice, m_pad, ldc));
    return DH_OK;
}

template <int CS>
int max_cluster_ctas() {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS, 1, 1);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nclusters = 0;
    cudaFuncSetAttribute(k_dino_gemm<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (cudaOccupancyMaxActiveClusters(&nclusters, k_dino_gemm<CS>, &cfg) != cudaSuccess) nclusters = 0;
    (void)cudaGetLastError();
    return nclusters * CS;
}

struct Plan { int m_tiles, n_pairs, nslices, kblocks, kb_per_slice, m_pad, ldc, cs, max_ctas; };

int make_plan(int32_t N, int32_t Fm, int64_t Kdim, Plan* pl) {
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    pl->m_tiles = (N + BM - 1) / BM;
    // clusters of 4 template tiles share the frame tile by multicast; fewer than 4 tiles: no cluster
    pl->cs = (pl->m_tiles >= 4) ? 4 : 1;
    if (const char* e = getenv("DH_DINO_CLUSTER")) {   // tuning knob: 1, 2, 4 or 8
        const int v = atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8) pl->cs = v;
    }
    pl->m_tiles = (pl->m_tiles + pl->cs - 1) / pl->cs * pl->cs;
    pl->n_pairs = (Fm + NT * BN - 1) / (NT * BN);
    pl->kblocks = (int)((Kdim + BK - 1) / BK);
    // clusters cannot use every SM (they are placed inside one GPC): ask the driver how many fit at once
    if (pl->cs > 1) {
        const int c = pl->cs == 2 ? max_cluster_ctas<2>() : (pl->cs == 4 ? max_cluster_ctas<4>() : max_cluster_ctas<8>());
        sms = c > 0 ? c : sms * 128 / 148;
    }
    pl->max_ctas = sms;
    const int tiles = pl->m_tiles * pl->n_pairs;
    int ns = sms / tiles;
    if (ns < 1) ns = 1;
    if (4 * tiles > sms) {
        // many output tiles (the reference's 6000 templates: 48 tiles): several waves.  Pick the K split whose last wave
        // is fullest, among splits that keep at least 64 k-blocks (8 pipeline refills) per CTA.
        double best = -1.0;
        for (int c = 1; c <= 64 && pl->kblocks / c >= 64; c++) {
            const double waves = (double)tiles * c / sms;
            const double fill = waves / (double)(long long)(waves + 0.999999);
            if (fill > best + 1e-9) { best = fill; ns = c; }
        }
    }
    if (ns > pl->kblocks) ns = pl->kblocks;
    if (ns > 64) ns = 64;
    pl->kb_per_slice = (pl->kblocks + ns - 1) / ns;
    pl->nslices = (pl->kblocks + pl->kb_per_slice - 1) / pl->kb_per_slice;
    pl->m_pad = pl->m_tiles * BM;
    pl->ldc = pl->n_pairs * NT * BN;
    return DH_OK;
}

}  // namespace

extern "C" {

int dh_dino_workspace_bytes(int32_t N, int32_t Fm, int64_t Kdim, int64_t* bytes) {
    DH_REQUIRE(bytes != nullptr && N > 0 && Fm > 0 && Kdim > 0, "bad arguments");
    Plan pl;
    make_plan(N, Fm, Kdim, &pl);
    // split-K partial tiles + a [Fm, N] score matrix (used when the caller does not ask for the scores)
    *bytes = (int64_t)pl.nslices * pl.m_pad * pl.ldc * 4 + (int64_t)Fm * N * 4;
    return DH_OK;
}

int dh_dino_plan_info(int32_t N, int32_t Fm, int64_t Kdim, int32_t* out8) {
    DH_REQUIRE(out8 != nullptr && N > 0 && Fm > 0 && Kdim > 0, "bad arguments");
    Plan pl;
    make_plan(N, Fm, Kdim, &pl);
    out8[0] = pl.m_tiles; out8[1] = pl.n_pairs; out8[2] = pl.nslices; out8[3] = pl.kblocks;
    out8[4] = pl.kb_per_slice; out8[5] = pl.cs; out8[6] = pl.max_ctas; out8[7] = pl.ldc;
    return DH_OK;
}

int dh_dino_topk(const void* templ_bf16, const void* frames_bf16, int32_t N, int32_t Fm, int64_t Kdim, int32_t k,
                 float* scores, float* topk_vals, int32_t* topk_idx, void* workspace, int64_t workspace_bytes,
                 void* stream) {
    DH_REQUIRE(templ_bf16 && frames_bf16 && topk_vals && topk_idx && workspace, "NULL pointer");
    DH_REQUIRE(N > 0 && Fm > 0 && Kdim > 0, "N, Fm, Kdim must be positive");
    DH_REQUIRE(k >= 1 && k <= MAX_TOPK && k <= N, "k must be in [1, min(N, 32)]");
    if (Kdim % 8 != 0) return dh::fail(DH_ERR_UNSUPPORTED, "Kdim must be a multiple of 8 (16-byte TMA row pitch)");
    if (((uintptr_t)templ_bf16 | (uintptr_t)frames_bf16 | (uintptr_t)workspace) & 15)
        return dh::fail(DH_ERR_INVALID, "banks and workspace must be 16-byte aligned");
    if ((size_t)N * sizeof(float) > 200 * 1024) return dh::fail(DH_ERR_UNSUPPORTED, "N > 51200 templates");
    Plan pl;
    make_plan(N, Fm, Kdim, &pl);
    const int64_t part_bytes = (int64_t)pl.nslices * pl.m_pad * pl.ldc * 4;
    DH_REQUIRE(workspace_bytes >= part_bytes + (int64_t)Fm * N * 4, "workspace too small");
    float* score_buf = scores != nullptr ? scores : reinterpret_cast<float*>((char*)workspace + part_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap map_a, map_b;
    int rc = make_map(&map_a, templ_bf16, N, Kdim, BM);
    if (rc) return rc;
    rc = make_map(&map_b, frames_bf16, Fm, Kdim, pl.cs > 1 ? NT * BN / pl.cs : BN);
    if (rc) return rc;
    float* ws = (float*)workspace;
    if (pl.cs == 1) {
        DH_CUDA(cudaFuncSetAttribute(k_dino_gemm<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        k_dino_gemm<1><<<dim3(pl.m_tiles, pl.n_pairs, pl.nslices), GEMM_THREADS, SMEM_BYTES, st>>>(
            map_a, map_b, ws, pl.kblocks, pl.kb_per_slice, pl.m_pad, pl.ldc);
    } else {
        const dim3 grid(pl.m_tiles, pl.n_pairs, pl.nslices);
        if (pl.cs == 2)      rc = launch_gemm_cluster<2>(grid, st, map_a, map_b, ws, pl.kblocks, pl.kb_per_slice, pl.m_pad, pl.ldc);
        else if (pl.cs == 4) rc = launch_gemm_cluster<4>(grid, st, map_a, map_b, ws, pl.kblocks, pl.kb_per_slice, pl.m_pad, pl.ldc);
        else                 rc = launch_gemm_cluster<8>(grid, st, map_a, map_b, ws, pl.kblocks, pl.kb_per_slice, pl.m_pad, pl.ldc);
        if (rc) return rc;
    }
    DH_LAUNCH_OK("k_dino_gemm");
    k_dino_reduce<<<dim3((N + 31) / 32, (Fm + 31) / 32), 256, 0, st>>>((const float*)workspace, pl.nslices, pl.m_pad,
                                                                       pl.ldc, N, Fm, score_buf);
    DH_LAUNCH_OK("k_dino_reduce");
    const size_t sm = (size_t)N * sizeof(float);
    DH_CUDA(cudaFuncSetAttribute(k_dino_topk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_dino_topk<<<Fm, 256, sm, st>>>(score_buf, N, k, topk_vals, topk_idx);
    DH_LAUNCH_OK("k_dino_topk");
    return DH_OK;
}

int dh_dino_prescale(const float* feats, const float* mask, int32_t n, int32_t P, int32_t D, void* out_bf16,
                     void* stream) {
    DH_REQUIRE(feats && out_bf16 && n > 0 && P > 0 && D > 0, "bad arguments");
    const long long patches = (long long)n * P;
    k_dino_prescale<<<(unsigned)((patches + 7) / 8), 256, 0, (cudaStream_t)stream>>>(feats, mask, n, P, D,
                                                                                    (__nv_bfloat16*)out_bf16);
    DH_LAUNCH_OK("k_dino_prescale");
    return DH_OK;
}

}  // extern "C"
