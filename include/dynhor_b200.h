/*
 * dynhor_b200.h -- C ABI of libdynhor_b200.so (hand-written sm_100a kernels for Dynhor/ObjTracker's joint
 * pose-optimisation hot path and DINO template matching).
 *
 * Conventions:
 *   - every entry point returns 0 on success, a negative dh_status on failure; dh_last_error() gives the
 *     thread-local message.  Nothing throws across the boundary.
 *   - all data pointers are DEVICE pointers unless the name ends in _host; the caller (PyTorch) owns every
 *     buffer, including scratch; sizes are given by the dh_*_bytes helpers / documented shapes.
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, no host synchronisation unless
 *     stated.  One host thread per GPU.
 *   - "fn" numbers faces after the renderer's fill_back doubling: fn in [0,F) = faces as given,
 *     fn in [F,2F) = the same faces with reversed winding.
 *
 * The reference has no FFI of its own (pure Python, SURVEY.md section 8b "Reference's own plugin API: none");
 * each entry point cites the Python interface it replaces, relative to /root/reference/ObjTracker.
 * The ctypes binding a maintainer would add is dynhor_b200/_lib.py.
 */
#ifndef DYNHOR_B200_H
#define DYNHOR_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum dh_status {
    DH_OK = 0,
    DH_ERR_INVALID = -1,   /* bad argument (shape / flag / null pointer)        */
    DH_ERR_CUDA = -2,      /* a CUDA runtime call or kernel launch failed        */
    DH_ERR_UNSUPPORTED = -3, /* valid request outside what the kernels implement */
    DH_ERR_NO_DEVICE = -4  /* no sm_100 device                                   */
} dh_status;

int dh_version(void);
const char* dh_last_error(void);
/* sizeof of the ABI structs as compiled (0 dh_sil, 1 dh_jointopt, 2 dh_corr): lets a binding check its layout */
int dh_struct_bytes(int32_t which);
/* sm count, compute capability of the current device */
int dh_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* tuning / test knobs.  knob 0: contributing pixels per frame up to which the fused backward uses its per-line
 * pixel lists (default 32248; frames above take the bitmap kernel; value < 0 restores the default).  Takes effect
 * for launches and graph captures made after the call. */
/* The batch schedule the fused backward uses for a chunk of n_items (face, winding) items (host-side copy of the
 * kernel's own function, for tests): starts_out[0 .. nb] = first item of every batch, starts_out[nb] = n_items;
 * returns nb (>= 0), or a negative status.  cap: entries of starts_out, at least 97.  No reference counterpart. */
int dh_bwd_schedule(int32_t n_items, int32_t* starts_out, int32_t cap);
int dh_tune_set(int32_t knob, int32_t value);

/* ------------------------------------------------------------------------------------------------
 * Silhouette renderer state.  Replaces nr.renderer.Renderer(image_size=S, K, R=I, t=0, orig_size,
 * anti_aliasing)(verts, faces, mode="silhouettes")   utils/losses.py:36-40,68;
 * pose_initializtion.py:98-105,146-147,160.
 * Raster resolution is = S*2 if aa else S; `is` must be a multiple of 32 and <= 512.
 * ------------------------------------------------------------------------------------------------ */
typedef struct dh_sil {
    int32_t B, V, F, S, aa;           /* frames, vertices, faces (before fill_back), output size, anti-aliasing */
    float near_, far_, eps, orig_size; /* renderer near/far planes, backward eps (1e-4), projection orig_size   */
    const int32_t* faces;              /* [F,3]   shared by all frames (run.py:158 stacks identical copies)      */
    const float* K;                    /* [B,3,3] ROI intrinsics in unit-image coordinates                       */
    /* scratch, caller-allocated (sizes from dh_sil_scratch_bytes, in this order) */
    float* proj;                       /* [B,V,4]  NDC u, v, depth z, pad                                        */
    int32_t* bin_count;                /* [B,nstrips,2] entries per (strip, winding group)                       */
    int32_t* bins;                     /* [B,nstrips,2F] given windings from the front, reversed from the back   */
    int32_t* fidx;                     /* [B,is,is] face index map (-1 none), rasteriser row order               */
    uint32_t* alpha_bits;              /* [B,is,is/32] coverage bitmap, rasteriser row order                     */
    uint32_t* pos_pool;                /* [B,S,ceil(S/32)] bitmap: dL/drend > 0                                  */
    uint32_t* neg_pool;                /* [B,S,ceil(S/32)] bitmap: dL/drend < 0                                  */
    float* gpool;                      /* reserved (16 bytes): dL/drend is rebuilt from the bitmaps (fused path) or read   */
                                       /* from the caller's grad_rend (dh_sil_backward); never materialised               */
    float* gmax;                       /* [B] max |dL/d(raster pixel)| per frame (scales the backward's fixed point) */
    uint32_t* owned;                   /* [B,ceil(2F/32)] bitmap: face fn owns at least one pixel of the frame       */
    uint32_t* negT;                    /* [B,is,is/32] column-major bitmap: pixel uncovered && dL/dpixel < 0         */
    int16_t* row_rng;                  /* [B,4,is] first / last set pixel of every row, then of every column       */
    uint16_t* neg_lists;               /* [B,2,32768] the same pixels as per-column / per-row lists (fused backward) */
} dh_sil;

/* bytes of each scratch array, out[13] in the struct's order (proj, bin_count, bins, fidx, alpha_bits, pos_pool,
 * neg_pool, gpool, gmax, owned, negT, row_rng, neg_lists) */
int dh_sil_scratch_bytes(int32_t B, int32_t V, int32_t F, int32_t S, int32_t aa, int64_t* out13);

/* rend[B,S,S] = silhouettes of camera-space vertices verts_cam[B,V,3]  (forward of the renderer call). */
int dh_sil_forward(const dh_sil* s, const float* verts_cam, float* rend, void* stream);
/* grad_verts[B,V,3] (overwritten) = pseudo-gradient of sum(grad_rend * rend) w.r.t. verts_cam; must follow a
 * dh_sil_forward on the same state (uses its fidx / alpha_bits).   autograd of utils/losses.py:68. */
int dh_sil_backward(const dh_sil* s, const float* verts_cam, const float* grad_rend, float* grad_verts,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * Small geometry ops (drop-in pieces of utils/geometry.py and utils/camera.py).
 * ------------------------------------------------------------------------------------------------ */
/* utils/geometry.py:7-25  rot6d [B,3,2] -> R [B,3,3] (columns b1,b2,b3) */
int dh_rot6d_to_matrix(const float* rot6d, float* R, int32_t B, void* stream);
/* utils/camera.py:179-207  out[B,V,3] = (|scale| * verts[V,3]) @ R[B] + T[B] */
int dh_transform_verts(const float* verts, const float* R, const float* T, const float* scale, float* out,
                       int32_t B, int32_t V, void* stream);
/* target masks f32 [B,S,S] in {-1,0,1} -> tri-state int8 + number of keep (>=0) pixels (device uint64) */
int dh_masks_prepare(const float* target_masks, int8_t* tri, unsigned long long* keep_count, int64_t n,
                     void* stream);
/* mesh moments in double: out[0..2] = sum v, out[3..11] = sum v v^T (row-major) */
int dh_mesh_moments(const float* verts, int32_t V, double* out12, void* stream);

/* ------------------------------------------------------------------------------------------------
 * [BUILDER-DEFINED] Dense-correspondence reprojection term.  BASELINE.json's north_star names "reprojection
 * residuals of the DKM dense correspondences"; the reference has NO such code (SURVEY.md section 0.3 -- only a
 * data-folder comment, README.md:43), so there is nothing to be at parity with except this build's own oracle
 * (oracle/corr_oracle.py).  Off by default: configs/custom_shoes.yaml has no lw_corr_obj, and the reference's
 * weighting rule (jointopt.py:147-150, loss name -> "lw" name) is what switches it on.
 *   record   = 6 floats: X[3] point in canonical mesh coordinates (lifted once from the source frame's pixel),
 *              t[2] target position in the frame's ROI unit-image coordinates (what K_roi maps into), w weight
 *   residual e = S * (K_roi (c.x/zc, c.y/zc, 1) - t),  c = (|s| X) R_b + T_b,  zc = c.z + 1e-9   [ROI pixels]
 *   loss_corr_obj = sum_{b,c} w * huber_delta(|e|) / sum_{b,c} w
 * One streaming pass over the records per iteration (24 B / record, HBM-bound): cp.async.bulk (TMA) tiles into a
 * shared-memory ring, 13 accumulators per thread, fixed-order CTA reduction -> partials (bit-reproducible).
 * ------------------------------------------------------------------------------------------------ */
typedef struct dh_corr {
    const float* records;          /* [B,C,6]; C must be even (16-byte tiles); pad with zero-weight records       */
    int32_t C;                     /* records per frame                                                          */
    int32_t nslots;                /* partial-sum rows per frame (dh_corr_plan)                                  */
    float delta;                   /* Huber threshold in ROI pixels                                              */
    float pad_;
    double w_sum;                  /* sum of the weights over ALL ranks                                          */
    double lw_corr;                /* loss weight, 0 disables the term                                           */
    float* partials;               /* [B,nslots,16] scratch: dT(3), X (x) dc (9), loss (1), pad(3)               */
    const double* w_sum_dev;       /* optional: the same sum in DEVICE memory, read by the kernels instead of     */
                                   /* w_sum (lets the host queue iterations before the records are uploaded)     */
} dh_corr;

/* Launch plan of the streaming kernel: a frame's ceil(C/1024) record tiles form `nslots` segments of 8 tiles; every
 * segment is summed by one CTA in one fixed order (so a frame's sums do not depend on B or on the grid), the
 * B*nslots segments are dealt to `grid` persistent CTAs (3 per SM) in contiguous, nearly equal ranges.
 * out3 = grid, nslots, tiles per frame.  sm_count 0 = ask the current device (148 if there is none). */
int dh_corr_plan(int32_t B, int32_t C, int32_t sm_count, int32_t* out3);
/* Un-normalised partial sums for the poses (Rmat [B,9], trans [B,3], scale [1]) and intrinsics K [B,9]:
 * partials[b,slot,0..2] = sum w dhuber/dc, [3..11] = sum X_i * (w dhuber/dc)_j, [12] = sum w huber; unused slots
 * are zeroed.  The caller sums over slots and applies lw_corr / w_sum; dL/dR = |s| * [3..11],
 * dL/d|s| = <R, [3..11]>.  partials: [B,nslots,16] with nslots from dh_corr_plan(B, C, 0). */
int dh_corr_eval(const float* records, int32_t B, int32_t C, const float* Rmat, const float* trans,
                 const float* scale, const float* K, int32_t S, float delta, float* partials, int32_t nslots,
                 void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused joint optimisation:  jointopt.py:144-160 (zero_grad, forward, weighting, backward, Adam step)
 * for the local frame range of one GPU.
 * ------------------------------------------------------------------------------------------------ */
typedef struct dh_jointopt {
    dh_sil sil;                    /* renderer state; sil.K, sil.faces as above                                 */
    const float* verts_og;         /* [V,3] canonical mesh                     jointopt.py:39                   */
    const int8_t* mask_tri;        /* [B,S,S] {-1 occluder, 0 background, 1 object}   jointopt.py:50-53         */
    float* rot6d;                  /* [B,3,2] parameters, updated in place     jointopt.py:37-38                */
    float* trans;                  /* [B,1,3] parameters, updated in place     jointopt.py:30-31                */
    float* scale;                  /* [1]     int_scales_object                jointopt.py:40-48                */
    float* adam_m_rot; float* adam_v_rot;      /* [B,6] each */
    float* adam_m_trans; float* adam_v_trans;  /* [B,3] each */
    float* adam_mv_scale;          /* [2] */
    int32_t* step;                 /* [1] device iteration counter (0 before the first step)                   */
    double* hist;                  /* [max_iters+1,4] per-iteration partial sums of this rank:                   */
                                   /*   loss_smooth_obj, loss_sil_obj, iou_object, loss_corr_obj; row max_iters  */
                                   /*   belongs to dh_jointopt_eval / dh_jointopt_grads                          */
    int32_t max_iters;
    /* neighbours' boundary poses for the smoothness term (frame-range sharding) */
    const float* halo_prev;        /* [9] rot6d(6)+trans(3) of global frame first-1, or NULL at the start        */
    const float* halo_next;        /* [9] of global frame last+1, or NULL at the end                            */
    /* Optional peer-to-peer halo (one process per GPU, CUDA IPC over NVLink).  When `mailbox` is set the halo_*
     * pointers are ignored: after its Adam step each rank stores its boundary poses straight into its neighbours'
     * mailboxes and raises a flag; the next iteration's pose kernel waits on the flag.  No host involvement.
     * Flags hold "ticks" = tick_base + iteration: a mailbox (and its IPC mappings) is allocated once per process
     * and reused by later runs, each run starting at a tick_base beyond every tick of the previous one. */
    float* mailbox;                /* [DH_MAILBOX_WORDS] this rank's mailbox (dh_dev_alloc), or NULL             */
    float* peer_prev;              /* mailbox of rank-1 mapped into this process (dh_ipc_open), NULL if none     */
    float* peer_next;              /* mailbox of rank+1, NULL if none                                           */
    int32_t tick_base;             /* tick of this run's iteration 0                                             */
    int32_t halo_timeout_ms;       /* a wait on a neighbour gives up after this long (0 = 60 s), records          */
                                   /* DH_STATUS_HALO_TIMEOUT in *status and lets the kernel finish               */
    int32_t* status;               /* [1] device word, 0 = ok; checked by the host after the run (may be NULL)   */
    /* shared object scale under frame sharding (jointopt.py:42-46: int_scales_object is ONE parameter of the
     * rigid Adam group, its gradient sums over all frames of all ranks) */
    int32_t rank, world;           /* this rank / number of ranks (world <= DH_MAX_RANKS); 0 / 1 when unsharded  */
    int32_t scale_mode;            /* DH_SCALE_LOCAL, DH_SCALE_P2P or DH_SCALE_DEFERRED                          */
    float* peers[16];              /* DH_SCALE_P2P: mailboxes of ALL ranks (peers[rank] = own mailbox)           */
    unsigned long long* scale_part; /* [2] DH_SCALE_DEFERRED: this rank's exact partial scale gradient (dh_fx128) */
    int32_t B_total;               /* frames in the whole sequence (all ranks)                                  */
    double keep_sum;               /* sum over ALL ranks of keep-mask pixels      utils/losses.py:71             */
    double lw_sil, lw_smooth;      /* loss weights, 0 disables a term             jointopt.py:81,86,147-150      */
    double lr;                     /* translations/scale lr; rotations use 10*lr  jointopt.py:135-141            */
    int32_t optimize_scale;        /* jointopt.py:42-46                                                         */
    const double* moments;         /* [12] from dh_mesh_moments                                                  */
    /* scratch */
    float* Rmat;                   /* [B,9]                                                                      */
    double* smooth_terms;          /* [B,16]: gT(3) gR(9) gs(1) pair_sse(1) pad(2)                               */
    int32_t* loss_counts;          /* [B,4]: 16*SSE, 4*inter, 4*union, pad                                       */
    float* partials;               /* [B,nchunks,16] per-CTA pose-gradient partial sums                          */
    double* frame_terms;           /* [B,8]: 16*SSE, iou, pair_sse, scale-grad, corr loss sum, off-screen, pad(2) */
    int32_t nchunks;
    dh_corr corr;                  /* optional correspondence term (corr.records == NULL or lw_corr == 0: off)   */
    /* Stage-1 mode: the silhouette term of the per-frame pose initialisation,
     * ObjTracker.coarse_forward + the loop of find_optimal_pose (pose_initializtion.py:143-155,346-360), batched over
     * frames x initialisations.  loss_mode DH_LOSS_STAGE1: per frame lw_sil * (1 - IoU(keep * silhouette, ref)) +
     * lw_offscreen * off-screen penalty of the projected vertices (:119-141), rendered WITHOUT anti-aliasing
     * (sil.aa must be 0, :98-105), no smoothness term, ONE Adam group (rotations and translations share lr, :346).
     * hist rows then hold: sum of the off-screen penalties, sum of (1 - IoU), mean IoU, 0. */
    int32_t loss_mode;             /* DH_LOSS_JOINT (jointopt.py) or DH_LOSS_STAGE1                              */
    double lw_offscreen;           /* 100000 in the reference (:154)                                             */
    float* offscreen;              /* [B,16] scratch: dT(3) dR(9) pad(1) penalty(1) pad(2), stage 1 only          */
    float* frame_coef;             /* [B,2]  scratch: per-frame dL/dpixel coefficients of the IoU loss            */
    /* Optional iteration clock: [max_iters,2] zero-initialised u64, per iteration (globaltimer ns) the moment this
     * rank's own work could start (after its waits on the neighbours) and the moment it ended.  Their difference is the
     * rank's compute time without the waiting: what a cost-weighted re-partition needs.  NULL = off. */
    unsigned long long* iter_ns;
} dh_jointopt;
#define DH_LOSS_JOINT 0
#define DH_LOSS_STAGE1 1

/* bytes of the dh_jointopt scratch arrays: out[5] = Rmat, smooth_terms, loss_counts, partials, frame_terms */
int dh_jointopt_scratch_bytes(int32_t B, int32_t nchunks, int64_t* out5);
/* recommended number of face chunks per frame for the backward kernel */
int dh_jointopt_default_chunks(int32_t B, int32_t F);
/* run n_iters fused iterations on `stream` (no host sync).  use_graph != 0: capture one iteration into a CUDA
 * graph (cached per plan address) and replay it. */
int dh_jointopt_run(const dh_jointopt* p, int32_t n_iters, int32_t use_graph, void* stream);
/* ONE iteration as two plain-stream halves (no graph): part 1 = pose preparation + the silhouette term's kernels
 * (projection ... backward), part 2 = correspondence kernel + pose update + bookkeeping.  The caller may make the
 * stream wait for the upload of the correspondence records between the halves, so that the first iteration's
 * silhouette work runs beside that copy.  Same arithmetic and results as dh_jointopt_run(p, 1, ...). */
int dh_jointopt_run_part(const dh_jointopt* p, int32_t part, void* stream);
/* forward only: losses of the current parameters into hist[max_iters] without touching parameters or step. */
int dh_jointopt_eval(const dh_jointopt* p, void* stream);
/* gradients of the weighted loss w.r.t. rot6d [B,6] and trans [B,3] (and scale [1]) for the current
 * parameters, without an optimiser step (parity tests; also the backward of Joint_Optimizer.forward). */
int dh_jointopt_grads(const dh_jointopt* p, float* grad_rot6d, float* grad_trans, float* grad_scale, void* stream);
/* run n_iters iterations eagerly with CUDA events around each kernel; ms_out_host[8] (HOST memory) = average
 * milliseconds of pose_prep, project, setup_bin, raster, backward (incl. its per-frame map kernel), pose_update,
 * finalize, corr.  Synchronises. */
int dh_jointopt_profile(const dh_jointopt* p, int32_t n_iters, float* ms_out_host, void* stream);
/* drop the cached graphs of this plan (NULL: all) */
int dh_jointopt_release(const dh_jointopt* p);

/* Mailbox layout (32-bit words; every rank owns one, written by its peers over NVLink):
 *   [0,128)    halo slots: slot(side, tick & 3) = 16 floats at (side*4 + (tick & 3))*16; side 0 = pose of global
 *              frame first-1, side 1 = pose of global frame last+1 (rot6d(6) + trans(3))
 *   [128,136)  halo flags (int32): tick the slot (side, tick & 3) is valid for
 *   [136,392)  scale slots: slot(tick & 3, rank) = 4 words (dh_fx128: hi int64, lo uint64) at 136 + ((tick&3)*16 + rank)*4
 *   [392,456)  scale flags (int32) at 392 + (tick&3)*16 + rank
 * Four slots per side and monotonically increasing ticks: a neighbour that is still publishing the last tick of the
 * previous run can never touch the slot the next run seeds (tick_base advances by n_iters + 2). */
#define DH_MAILBOX_WORDS 512
#define DH_MAX_RANKS 16
#define DH_SCALE_LOCAL 0     /* single rank: k_finalize applies the Adam step to the scale                         */
#define DH_SCALE_P2P 1       /* every rank stores its exact partial gradient into all mailboxes, sums them in     */
                             /* rank order (bit-identical on every rank) and applies the step; no host work       */
#define DH_SCALE_DEFERRED 2  /* k_finalize only stores the partial (scale_part); the host gathers the partials of */
                             /* all ranks and calls dh_scale_apply (the "nccl" / gloo path)                       */
#define DH_STATUS_HALO_TIMEOUT 1
/* Exact order-independent sum of doubles: value = hi + lo * 2^-64 (two's complement, floor).  The scale gradient is
 * accumulated in this form so that a frame-sharded run gives the same bits as a single-GPU run whatever the
 * partition.  parts: [world][2] u64 (hi, lo) in rank order, HOST or DEVICE memory as stated. */
/* Adam step of the shared scale from the gathered partial gradients (device memory, [world][2]); uses *p->step as
 * the 1-based step number (call it after the iteration's dh_jointopt_run). */
int dh_scale_apply(const dh_jointopt* p, const unsigned long long* parts_dev, int32_t world, void* stream);
/* Times the heavy kernels (projection, binning, raster, backward) of one iteration separately for `nblocks`
 * contiguous, equal blocks of this plan's frames, without touching parameters or optimiser state:
 * ms_out_host[0..nblocks) = milliseconds per block, ms_out_host[nblocks] = the correspondence kernel over all
 * frames (0 when off).  Used to cut a sequence into frame ranges of equal COST (sharding.balanced_bounds).
 * Synchronises the stream. */
int dh_jointopt_probe(const dh_jointopt* p, int32_t nblocks, float* ms_out_host, void* stream);
/* device memory that can be exported to the other ranks of the box (a dedicated cudaMalloc, zero-filled) */
int dh_dev_alloc(void** ptr, int64_t bytes);
int dh_dev_free(void* ptr);
int dh_memcpy_d2d(void* dst, const void* src, int64_t bytes, void* stream);
/* n host rows of row_bytes each (src_rows_host: n HOST pointers, pinned memory for an asynchronous copy) -> the
 * contiguous device rows dst + i * row_bytes, as ONE batched driver call where the runtime offers it.  Replaces the
 * per-frame `.cuda()` calls of the reference's loaders (jointopt.py:104-123 concatenates per-frame tensors that
 * pose_initializtion.py:460-471 left on the device one by one). */
int dh_upload_rows(void* dst, const void* const* src_rows_host, int64_t row_bytes, int32_t n, void* stream);
/* CUDA IPC: export / open / close a dh_dev_alloc allocation (handle = 64 bytes of HOST memory) */
int dh_ipc_export(const void* ptr, void* handle64_host);
int dh_ipc_open(const void* handle64_host, void** ptr);
int dh_ipc_close(void* ptr);

/* torch.optim.Adam single step on a flat fp32 tensor (jointopt.py:135-141,160), t = 1-based step number */
int dh_adam_step(float* param, const float* grad, float* m, float* v, int64_t n, double lr, int32_t t,
                 void* stream);

/* ------------------------------------------------------------------------------------------------
 * DINO template matching:  pose_initializtion.py:295-296 (aligned-patch masked cosine) + :299,309
 * (argmax / topk).  Inputs are pre-scaled bf16 banks so that score[n,f] = <templ[n,:], frames[f,:]>:
 *   templ [N,Kdim]  = r[n,p,:] / |r[n,p,:]|
 *   frames[Fm,Kdim] = m_f[p] * g[f,p,:] / |g[f,p,:]| / sum_p m_f[p]          (Kdim = P*D)
 * Output: scores fp32 [Fm,N] (optional, may be NULL), top-k values [Fm,k] and indices [Fm,k] (largest first,
 * lowest index first on exact ties).
 * ------------------------------------------------------------------------------------------------ */
int dh_dino_workspace_bytes(int32_t N, int32_t Fm, int64_t Kdim, int64_t* bytes);
/* launch plan of dh_dino_topk: out8 = template tiles, frame tile pairs, K slices, k-blocks, k-blocks per slice,
 * cluster size, co-resident CTAs the plan was sized for, workspace row pitch */
int dh_dino_plan_info(int32_t N, int32_t Fm, int64_t Kdim, int32_t* out8);
int dh_dino_topk(const void* templ_bf16, const void* frames_bf16, int32_t N, int32_t Fm, int64_t Kdim, int32_t k,
                 float* scores, float* topk_vals, int32_t* topk_idx, void* workspace, int64_t workspace_bytes,
                 void* stream);
/* builds the pre-scaled bf16 banks from fp32 features [n,P,D] (+ optional mask [n,P], NULL = all ones) */
int dh_dino_prescale(const float* feats, const float* mask, int32_t n, int32_t P, int32_t D, void* out_bf16,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * ROI preprocessing: the target masks the joint optimisation consumes.
 * Replaces the per-frame CPU loop ObjTracker/run.py:26-72 (process_input) and its helpers utils/bbox.py:8-36
 * (crop_and_resize = detectron2 ROIAlign((S,S), 1.0, 0, aligned=True)), utils/bbox.py:73-117 (square box, box
 * modes) and utils/maskutils.py:8-30 (add_occlusions), for all B frames in three launches.
 *   obj_bits, hand_bits  [B,H,W] u8, 1 where the SAM mask == 255 (hand_bits may be NULL: no occluder)
 *   images_hwc           [B,H,W,3] u8 (NULL with crop_image NULL: masks only)
 *   bounds               [B,4] i32 scratch -> min_row, max_row, min_col, max_col (max_row < 0: empty object mask;
 *                        the reference raises there, the caller must check)
 *   bbox, square_bbox    [B,4] f32 xywh  (run.py:41-43)
 *   crop_mask            [B,S,S] u8      (run.py:47)        target [B,S,S] f32 in {1, 0, -1} (run.py:66-68)
 *   target_tri           [B,S,S] i8 or NULL (the same values, the fused iteration's mask format)
 *   crop_image           [B,3,S,S] f32 or NULL, white outside the object mask (run.py:49-51)
 * ------------------------------------------------------------------------------------------------ */
int dh_roi_process(const uint8_t* obj_bits, const uint8_t* hand_bits, const uint8_t* images_hwc, int32_t B,
                   int32_t H, int32_t W, int32_t S, float pad, float expansion, int32_t* bounds, float* bbox,
                   float* square_bbox, uint8_t* crop_mask, float* target, int8_t* target_tri, float* crop_image,
                   void* stream);

/* The same crops for rendered TEMPLATE views: pose_initializtion.py:188-246 (compute_prior_features) builds, per view,
 * the tight box of the rendering's alpha == 1 mask (+5 px, clamped to the render size), the square box x1.3, and
 * ROIAlign crops of the mask, the float RGB rendering (white outside the mask, :215) and the depth map -- one view at
 * a time, before a batch-1 DINOv2 forward.  Here: all views of a batch in three launches.
 *   obj_bits [B,H,W] u8 (rendering alpha == 1), images [B,H,W,pitch] f32 (RGB in channels 0..2, pitch 4 for RGBA),
 *   depth [B,H,W] f32 or NULL;  crop_image [B,3,S,S] f32, crop_depth [B,S,S] f32 or NULL; the rest as above. */
int dh_roi_process_f32(const uint8_t* obj_bits, const float* images, int32_t pitch, const float* depth, int32_t B,
                       int32_t H, int32_t W, int32_t S, float pad, float expansion, int32_t* bounds, float* bbox,
                       float* square_bbox, uint8_t* crop_mask, float* crop_image, float* crop_depth, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DYNHOR_B200_H */
