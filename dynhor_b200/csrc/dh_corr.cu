// dh_corr.cu -- [BUILDER-DEFINED] dense-correspondence reprojection term of the joint optimisation.
//
// BASELINE.json's north_star lists "reprojection residuals of the DKM dense correspondences" on the hot path; the
// reference itself has no such code, so the definition is this build's (include/
// dynhor_b200.h, oracle/corr_oracle.py) and the term is off unless the caller passes records and lw_corr_obj.
//
// k_corr is the one purely HBM-bound kernel of the iteration: 24 bytes per record are read once, ~60 flops each.
//   * persistent grid (3 CTAs per SM); the B * ceil(C/1024) record tiles are cut into gridDim.x contiguous ranges of
//     (nearly) equal length whose ends are moved to SEGMENT boundaries (a segment = 8 consecutive tiles of one frame);
//     a segment is therefore always summed by one CTA, in one fixed order, whatever B and the grid are: a frame's
//     result does not depend on how many frames share the launch (frame-sharded runs reproduce single-GPU bits);
//   * tiles (1024 records = 24 KB) arrive through cp.async.bulk (TMA, no tensor map needed for a 1-D copy) into a
//     3-stage shared-memory ring guarded by mbarriers: one thread issues, 256 threads consume;
//   * shared-memory reads are 8-byte accesses at an odd stride (3 float2 per record): conflict-free;
//   * 13 accumulators per thread, reduced per segment with warp shuffles in a fixed order and stored to the
//     segment's slot of the frame: no atomics, bit-reproducible.
#include "dh_common.h"
#include "dh_core.h"

namespace {

using namespace dh;

constexpr int kCorrThreads = 256;
constexpr int kCorrTile = 1024;                       // records per stage
constexpr int kSegTiles = 8;                          // tiles per segment (the unit of summation)
#ifndef DH_CORR_STAGES
#define DH_CORR_STAGES 3
#endif
#ifndef DH_CORR_CTAS
#define DH_CORR_CTAS 3
#endif
constexpr int kCorrStages = DH_CORR_STAGES;
constexpr int kRecBytes = 24;
constexpr int kStageBytes = kCorrTile * kRecBytes;    // 24 KB
constexpr int kCorrCtasPerSm = DH_CORR_CTAS;          // 3 x 72 KB of shared memory per SM

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {  // bounded: a protocol error traps
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

struct CorrPlan { int grid, nslots, tpf; };

__host__ __device__ inline CorrPlan corr_plan(int B, int C, int sms) {
    CorrPlan p;
    p.tpf = (C + kCorrTile - 1) / kCorrTile;
    p.nslots = (p.tpf + kSegTiles - 1) / kSegTiles;   // segments per frame
    const long long segs = (long long)B * p.nslots;
    long long g = (long long)sms * kCorrCtasPerSm;
    if (g > segs) g = segs;
    if (g < 1) g = 1;
    p.grid = (int)g;
    return p;
}

// First tile of CTA c's range: the cut c * T / G moved forward to the next segment start.
__host__ __device__ inline long long corr_cut(long long c, long long T, long long G, int tpf) {
    const long long t = c * T / G;
    const long long b = t / tpf;
    const int k = (int)(t - b * tpf);
    const int ks = ((k + kSegTiles - 1) / kSegTiles) * kSegTiles;   // segment starts inside a frame: 0, 8, 16, ...
    return b * tpf + (ks < tpf ? ks : tpf);
}

// tile t of the flattened (frame, tile-in-frame) list -> records pointer and count
__device__ __forceinline__ void tile_span(long long t, int tpf, int C, const float* records, const float** src,
                                          int* nrec) {
    const int b = (int)(t / tpf), k = (int)(t - (long long)b * tpf);
    *src = records + ((size_t)b * C + (size_t)k * kCorrTile) * 6;
    *nrec = min(kCorrTile, C - k * kCorrTile);
}

// ---- packed fp32 (sm_100 FFMA2 / FMUL2 / FADD2): one instruction works on two records
__device__ __forceinline__ float2 bc(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_fast(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Per-frame constants of the pair kernel, every value duplicated into both halves of a float2.
struct CorrPose {
    float2 Rs[9];   // |s| R
    float2 T[3];
    float2 SK[6];   // S * K (first two rows)
    float2 nS;      // -S
};

// dh_core.h::corr_record for two records at once (lanes .x / .y), approximate reciprocal / rsqrt (relative error
// 2^-22, far inside the 1e-4 / 1e-3 parity bars) and a branch-free Huber.  acc: 13 float2 accumulators.
__device__ __forceinline__ void corr_pair(const float4 v0, const float4 v1, const float4 v2, const CorrPose& P,
                                          float delta, float2* acc) {
    const float2 X0 = make_float2(v0.x, v1.z), X1 = make_float2(v0.y, v1.w), X2 = make_float2(v0.z, v2.x);
    const float2 tu = make_float2(v0.w, v2.y), tv = make_float2(v1.x, v2.z), w = make_float2(v1.y, v2.w);
    const float2 cx = fma2(X0, P.Rs[0], fma2(X1, P.Rs[3], fma2(X2, P.Rs[6], P.T[0])));
    const float2 cy = fma2(X0, P.Rs[1], fma2(X1, P.Rs[4], fma2(X2, P.Rs[7], P.T[1])));
    const float2 cz = fma2(X0, P.Rs[2], fma2(X1, P.Rs[5], fma2(X2, P.Rs[8], P.T[2])));
    const float2 iz = make_float2(rcp_fast(cz.x + 1e-9f), rcp_fast(cz.y + 1e-9f));
    const float2 x_ = mul2(cx, iz), y_ = mul2(cy, iz);
    const float2 eu = fma2(x_, P.SK[0], fma2(y_, P.SK[1], fma2(tu, P.nS, P.SK[2])));
    const float2 ev = fma2(x_, P.SK[3], fma2(y_, P.SK[4], fma2(tv, P.nS, P.SK[5])));
    const float2 r2 = fma2(eu, eu, mul2(ev, ev));
    const float d2 = delta * delta, hd = 0.5f * delta;
    const float rsx = rsqrt_fast(fmaxf(r2.x, 1e-30f)), rsy = rsqrt_fast(fmaxf(r2.y, 1e-30f));
    const bool qx = r2.x <= d2, qy = r2.y <= d2;
    const float2 f = make_float2(qx ? 1.0f : delta * rsx, qy ? 1.0f : delta * rsy);
    const float2 rho = make_float2(qx ? 0.5f * r2.x : delta * (r2.x * rsx - hd),
                                   qy ? 0.5f * r2.y : delta * (r2.y * rsy - hd));
    const float2 wf = mul2(w, f);
    const float2 gu = mul2(wf, eu), gv = mul2(wf, ev);
    const float2 gx_ = fma2(gu, P.SK[0], mul2(gv, P.SK[3])), gy_ = fma2(gu, P.SK[1], mul2(gv, P.SK[4]));
    const float2 g0 = mul2(gx_, iz), g1 = mul2(gy_, iz);
    const float2 t = fma2(gx_, x_, mul2(gy_, y_));
    const float2 g2 = mul2(t, make_float2(-iz.x, -iz.y));
    acc[0] = add2(acc[0], g0); acc[1] = add2(acc[1], g1); acc[2] = add2(acc[2], g2);
    acc[3] = fma2(X0, g0, acc[3]); acc[4] = fma2(X0, g1, acc[4]); acc[5] = fma2(X0, g2, acc[5]);
    acc[6] = fma2(X1, g0, acc[6]); acc[7] = fma2(X1, g1, acc[7]); acc[8] = fma2(X1, g2, acc[8]);
    acc[9] = fma2(X2, g0, acc[9]); acc[10] = fma2(X2, g1, acc[10]); acc[11] = fma2(X2, g2, acc[11]);
    acc[12] = fma2(w, rho, acc[12]);
}

__global__ void __launch_bounds__(kCorrThreads, kCorrCtasPerSm)
k_corr(const float* __restrict__ records, int B, int C, int tpf, int nslots, const float* __restrict__ Rmat,
       const float* __restrict__ trans, const float* __restrict__ scale, const float* __restrict__ K, float S,
       float delta, float* __restrict__ partials) {
    extern __shared__ __align__(128) unsigned char ring[];          // kCorrStages x kStageBytes
    __shared__ __align__(8) unsigned long long bars[kCorrStages];
    __shared__ float red[kCorrThreads / 32][13];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long T = (long long)B * tpf, G = gridDim.x;
    const long long t0 = corr_cut(blockIdx.x, T, G, tpf), t1 = corr_cut((long long)blockIdx.x + 1, T, G, tpf);
    const int ntiles = (int)(t1 - t0);
    if (tid == 0) {
        for (int i = 0; i < kCorrStages; i++) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int i = 0; i < kCorrStages && i < ntiles; i++) {
            const float* src; int nrec;
            tile_span(t0 + i, tpf, C, records, &src, &nrec);
            mbar_expect_tx(smem_u32(&bars[i]), (uint32_t)nrec * kRecBytes);
            bulk_load(smem_u32(ring + i * kStageBytes), src, (uint32_t)nrec * kRecBytes, smem_u32(&bars[i]));
        }
    }
    const float s_abs = fabsf(scale[0]);
    float2 acc[13];
    CorrPose P;
    P.nS = bc(-S);
    int cur_b = -1;
    for (int i = 0; i < ntiles; i++) {
        const long long t = t0 + i;
        const int b = (int)(t / tpf), k = (int)(t - (long long)b * tpf);
        if (k % kSegTiles == 0) {
#pragma unroll
            for (int j = 0; j < 13; j++) acc[j] = bc(0.0f);
        }
        if (b != cur_b) {  // first tile of a frame: its pose and intrinsics
            cur_b = b;
#pragma unroll
            for (int j = 0; j < 9; j++) P.Rs[j] = bc(s_abs * Rmat[9 * b + j]);
#pragma unroll
            for (int j = 0; j < 3; j++) P.T[j] = bc(trans[3 * b + j]);
#pragma unroll
            for (int j = 0; j < 6; j++) P.SK[j] = bc(S * K[9 * b + j]);
        }
        const int stage = i % kCorrStages;
        mbar_wait(smem_u32(&bars[stage]), (uint32_t)(i / kCorrStages) & 1u);
        const int npairs = min(kCorrTile, C - k * kCorrTile) >> 1;       // C is even
        // a pair of records = 48 bytes = three 16-byte shared-memory loads; lanes 48 B apart are conflict-free
        const float4* tile = reinterpret_cast<const float4*>(ring + stage * kStageBytes);
#pragma unroll
        for (int j = 0; j < kCorrTile / 2 / kCorrThreads; j++) {
            const int p = tid + j * kCorrThreads;
            if (p < npairs) corr_pair(tile[3 * p], tile[3 * p + 1], tile[3 * p + 2], P, delta, acc);
        }
        __syncthreads();  // every thread is done with this stage: it may be refilled
        if (tid == 0 && i + kCorrStages < ntiles) {
            const float* src; int nr;
            tile_span(t + kCorrStages, tpf, C, records, &src, &nr);
            mbar_expect_tx(smem_u32(&bars[stage]), (uint32_t)nr * kRecBytes);
            bulk_load(smem_u32(ring + stage * kStageBytes), src, (uint32_t)nr * kRecBytes, smem_u32(&bars[stage]));
        }
        // last tile of a segment: reduce and store its sums into the segment's slot of frame b
        if (k % kSegTiles == kSegTiles - 1 || k == tpf - 1) {
#pragma unroll
            for (int j = 0; j < 13; j++) {
                float v = acc[j].x + acc[j].y;
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) red[warp][j] = v;
            }
            __syncthreads();
            if (tid < 16) {
                float v = 0.0f;
                if (tid < 13)
                    for (int w = 0; w < kCorrThreads / 32; w++) v += red[w][tid];
                partials[((size_t)b * nslots + k / kSegTiles) * 16 + tid] = v;
            }
            __syncthreads();
        }
    }
}

int device_sms() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
        cudaGetLastError();
        return 148;
    }
    return sms;
}

}  // namespace

namespace dh {

int launch_corr(const float* records, int B, int C, const float* Rmat, const float* trans, const float* scale,
                const float* K, int S, float delta, float* partials, int nslots, cudaStream_t st) {
    DH_REQUIRE(records && Rmat && trans && scale && K && partials, "corr: NULL pointer");
    DH_REQUIRE(B > 0 && C > 0 && S > 0, "corr: B, C, S must be positive");
    DH_REQUIRE(C % 2 == 0, "corr: C must be even (16-byte bulk-copy tiles); pad with a zero-weight record");
    DH_REQUIRE(((uintptr_t)records & 15u) == 0, "corr: records must be 16-byte aligned");
    const CorrPlan pl = corr_plan(B, C, device_sms());
    DH_REQUIRE(nslots >= pl.nslots, "corr: partials hold %d slots per frame, the plan needs %d", nslots, pl.nslots);
    static bool attr_set = false;
    const int smem = kCorrStages * kStageBytes;
    if (!attr_set) {
        DH_CUDA(cudaFuncSetAttribute(k_corr, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    DH_CUDA(cudaMemsetAsync(partials, 0, (size_t)B * nslots * 16 * sizeof(float), st));
    k_corr<<<pl.grid, kCorrThreads, smem, st>>>(records, B, C, pl.tpf, nslots, Rmat, trans, scale, K, (float)S, delta,
                                              partials);
    DH_LAUNCH_OK("k_corr");
    return DH_OK;
}

}  // namespace dh

extern "C" {

int dh_corr_plan(int32_t B, int32_t C, int32_t sm_count, int32_t* out3) {
    DH_REQUIRE(out3 != nullptr && B > 0 && C > 0, "bad arguments");
    const CorrPlan pl = corr_plan(B, C, sm_count > 0 ? sm_count : device_sms());
    out3[0] = pl.grid; out3[1] = pl.nslots; out3[2] = pl.tpf;
    return DH_OK;
}

int dh_corr_eval(const float* records, int32_t B, int32_t C, const float* Rmat, const float* trans,
                 const float* scale, const float* K, int32_t S, float delta, float* partials, int32_t nslots,
                 void* stream) {
    return dh::launch_corr(records, B, C, Rmat, trans, scale, K, S, delta, partials, nslots, (cudaStream_t)stream);
}

}  // extern "C"
