"""DINO template matching of ObjTracker's view selection on the B200 tensor cores.

Replaces the two steps of pose_initializtion.py:286-321 that are data-parallel over templates and frames:
    dino_cos = (cos_mask * sum(gt_feat * render_feats, -1) / (|gt_feat| |render_feats| + 1e-6)).sum(1) / cos_mask.sum(1)
                                                                              pose_initializtion.py:295-296
    torch.argmax(dino_cos) / torch.topk(dino_cos, k, largest=True)            pose_initializtion.py:299,309
The scores depend only on (frame features, template bank), not on the previous frame, so all frames are scored in
one GEMM; the sequential candidate gating (:300-321) stays on the host in `select_view`.
Tolerance: the banks are bf16 and fold the reference's `+ 1e-6` away, so scores agree with the fp32 expression to about
2e-3 absolute; top-k is index-exact where consecutive scores differ by more than that.  `rescore_topk_fp32` restores
the reference's order (and exact scores) among the k candidates of a frame when that matters.
Everything here needs CUDA tensors and the native library; there is no CPU fallback.
"""
import ctypes

import torch

from . import _lib
from .geometry import rotation_angle_difference


def build_bank(feats, mask=None):
    """feats [n,P,D] fp32 (CUDA) -> bf16 [n, P*D]: every patch vector divided by its norm (the reference
    normalises with F.normalize and divides by the norms again, :226,293,296) and, for frames, weighted by
    mask[n,p] / sum_p mask[n,:] (the nearest-resized foreground mask, :290,294).  Templates: mask=None."""
    if not feats.is_cuda:
        raise _lib.DynhorError("dynhor_b200.dino_match needs CUDA tensors (no CPU fallback)")
    f = feats.detach().contiguous().float()
    n, P, D = f.shape
    if (P * D) % 8 != 0:
        raise ValueError("P*D must be a multiple of 8")
    m = None
    if mask is not None:
        m = mask.detach().reshape(n, P).contiguous().float()
        if bool((m.sum(1) <= 0).any()):
            raise ValueError("every frame mask needs at least one foreground patch")
    out = torch.empty(n, P * D, dtype=torch.bfloat16, device=f.device)
    _lib.check(_lib.load().dh_dino_prescale(_lib.ptr(f), _lib.ptr(m), n, P, D, _lib.ptr(out), _lib.stream_ptr()),
               "dh_dino_prescale")
    return out


def dino_cos_topk(frame_bank, templ_bank, k, return_scores=True):
    """frame_bank [Fm,K] bf16, templ_bank [N,K] bf16 (from build_bank) -> (dino_cos [Fm,N] fp32 or None,
    topk values [Fm,k], topk indices [Fm,k] int64), largest first like torch.topk(largest=True).  With the scores:
    the torch.library op dynhor::dino_topk (dynhor_b200/ops.py)."""
    if not (frame_bank.is_cuda and templ_bank.is_cuda):
        raise _lib.DynhorError("dynhor_b200.dino_match needs CUDA tensors (no CPU fallback)")
    if return_scores:
        from . import ops  # noqa: F401  (registers torch.ops.dynhor.*)
        return torch.ops.dynhor.dino_topk(frame_bank, templ_bank, int(k))
    return _dino_cos_topk(frame_bank, templ_bank, k, return_scores=False)


def _dino_cos_topk(frame_bank, templ_bank, k, return_scores=True):
    assert frame_bank.dtype == torch.bfloat16 and templ_bank.dtype == torch.bfloat16
    assert frame_bank.is_contiguous() and templ_bank.is_contiguous()
    Fm, K = frame_bank.shape
    N, K2 = templ_bank.shape
    assert K == K2, "banks disagree on P*D"
    lib = _lib.load()
    nbytes = ctypes.c_int64()
    _lib.check(lib.dh_dino_workspace_bytes(N, Fm, K, ctypes.byref(nbytes)), "dh_dino_workspace_bytes")
    dev = frame_bank.device
    ws = torch.empty(max(int(nbytes.value), 16), dtype=torch.uint8, device=dev)
    scores = torch.empty(Fm, N, dtype=torch.float32, device=dev) if return_scores else None
    vals = torch.empty(Fm, k, dtype=torch.float32, device=dev)
    idx = torch.empty(Fm, k, dtype=torch.int32, device=dev)
    _lib.check(lib.dh_dino_topk(_lib.ptr(templ_bank), _lib.ptr(frame_bank), N, Fm, K, int(k), _lib.ptr(scores),
                                _lib.ptr(vals), _lib.ptr(idx), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
               "dh_dino_topk")
    return scores, vals, idx.long()


def rescore_topk_fp32(frame_feats, frame_masks, templ_feats, topk_idx):
    """The bf16 banks drop the reference's +1e-6 in the denominator and round every product to 8 mantissa bits: scores
    agree to ~2e-3, which can reorder candidates whose fp32 scores are closer than that (and `select_view` compares
    scores against max - std).  This re-scores only the k candidates of every frame with the reference expression
    itself (pose_initializtion.py:295-296, fp32, eps 1e-6) and re-sorts them -- k x P x D work per frame instead of
    N x P x D.  frame_feats [Fm,P,D] and frame_masks [Fm,P] on the GPU; templ_feats [N,P,D] fp32 on any device (only
    the candidate rows are fetched); topk_idx [Fm,k].  Returns (values [Fm,k] fp32, indices [Fm,k]) best first."""
    Fm, k = topk_idx.shape
    dev = frame_feats.device
    rows = templ_feats[topk_idx.reshape(-1).to(templ_feats.device)].to(dev, non_blocking=True).float()
    rows = rows.reshape(Fm, k, *rows.shape[1:])                                          # [Fm,k,P,D]
    g = frame_feats.float()[:, None]                                                    # [Fm,1,P,D]
    m = frame_masks.float()[:, None]                                                    # [Fm,1,P]
    cos = (m * (g * rows).sum(-1) / (g.norm(dim=-1) * rows.norm(dim=-1) + 1e-6)).sum(-1) / m.sum(-1)
    order = torch.argsort(topk_idx, dim=1, stable=True)           # lowest index first, then a stable sort by score:
    cos, idx = torch.gather(cos, 1, order), torch.gather(topk_idx, 1, order)            # ties keep the lowest index
    order = torch.argsort(cos, dim=1, descending=True, stable=True)
    return torch.gather(cos, 1, order), torch.gather(idx, 1, order)


def merge_topk(vals_by_rank, idx_by_rank, offsets, k):
    """Top-k over templates sharded across ranks: every rank scored its own slice of the bank (pose_initializtion.py:
    295-296 is independent per template) and kept its k best per frame; the world x k candidates of a frame are merged
    here -- largest value first, lowest GLOBAL template index on ties, the kernel's own order.
        vals_by_rank / idx_by_rank: [world, Fm, k] (idx local to the rank's slice); offsets[r] = first template of
        rank r.  Returns (values [Fm,k], global indices [Fm,k] int64)."""
    W, Fm, kk = vals_by_rank.shape
    off = torch.as_tensor(offsets, dtype=torch.int64, device=idx_by_rank.device).reshape(W, 1, 1)
    v = vals_by_rank.permute(1, 0, 2).reshape(Fm, W * kk)
    g = (idx_by_rank.to(torch.int64) + off).permute(1, 0, 2).reshape(Fm, W * kk)
    order = torch.argsort(g, dim=1, stable=True)                 # by index first ...
    v, g = torch.gather(v, 1, order), torch.gather(g, 1, order)
    order = torch.argsort(v, dim=1, descending=True, stable=True)   # ... then a stable sort by value keeps it on ties
    return torch.gather(v, 1, order)[:, :k], torch.gather(g, 1, order)[:, :k]


def dino_topk_sharded(frame_bank, templ_bank_local, k, rank, world, n_local_by_rank, group=None):
    """The reference-sized bank (6000 templates x 1369 x 768 bf16 = 12.6 GB) split over the GPUs of the box: this
    rank scores `templ_bank_local` (its contiguous slice), then one all_gather of the [Fm,k] lists and `merge_topk`.
    No scores matrix is returned (it would be [Fm, N_total] gathered from every rank)."""
    import torch.distributed as dist
    _, vals, idx = dino_cos_topk(frame_bank, templ_bank_local, min(k, templ_bank_local.shape[0]), return_scores=False)
    if vals.shape[1] < k:   # a slice with fewer than k templates: pad with -inf candidates
        pad = k - vals.shape[1]
        vals = torch.cat([vals, vals.new_full((vals.shape[0], pad), float("-inf"))], 1)
        idx = torch.cat([idx, idx.new_zeros((idx.shape[0], pad))], 1)
    if world == 1:
        return vals, idx
    vs = [torch.empty_like(vals) for _ in range(world)]
    is_ = [torch.empty_like(idx) for _ in range(world)]
    dist.all_gather(vs, vals.contiguous(), group=group)
    dist.all_gather(is_, idx.contiguous(), group=group)
    offsets = [int(sum(n_local_by_rank[:r])) for r in range(world)]
    return merge_topk(torch.stack(vs), torch.stack(is_), offsets, k)


def select_view(dino_cos, topk_indices, render_rotations, rotations_init=None, former_max_idx=None, use_former=True):
    """The sequential part of the view selection (pose_initializtion.py:298-321) for ONE frame, on the scores and
    top-k indices the kernels produced for all frames at once.
        dino_cos [N]; topk_indices [>= 10], best first; render_rotations [N,3,3];
        rotations_init [1,3,3] = the previous frame's optimised rotation (None on the first frame);
        former_max_idx = the template the previous frame picked (-1: it kept its predecessor's rotation).
    Returns the template index, or -1 = "keep the previous rotation" (:323-325).
    Among the 5 best-scoring views (10 after a frame without a pick) the one closest in angle to the previous pose
    wins unless it is more than 85 degrees from that pose or from the previous pick; otherwise the view nearest to
    the previous pose is taken if it is within 15 degrees, within 30 degrees of the previous pick and scores no worse
    than one standard deviation below the best score."""
    if not use_former or rotations_init is None:
        return int(topk_indices[0])
    views = render_rotations.transpose(1, 2)
    to_prev = rotation_angle_difference(rotations_init, views)           # every template against the previous pose
    has_former = former_max_idx != -1
    to_former = rotation_angle_difference(views[former_max_idx:former_max_idx + 1], views) if has_former \
        else torch.zeros_like(to_prev)
    shortlist = topk_indices[:5 if has_former else 10]
    pick = int(shortlist[torch.argmin(to_prev[shortlist])])
    if not (to_prev[pick] > 85.0 or to_former[pick] > 85.0):
        return pick
    near = int(torch.argmin(to_prev))
    if not (to_prev[near] < 15.0):
        return -1
    if has_former and to_former[near].item() > 30.0:
        return -1
    if dino_cos[near] < dino_cos.max() - dino_cos.std():
        return -1
    return near
