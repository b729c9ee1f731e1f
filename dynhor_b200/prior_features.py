"""Template feature bank of ObjTracker's view selection, built in batches and kept on the GPU.

Replaces compute_prior_features (pose_initializtion.py:188-246), which walks the rendered template views ONE AT A TIME
-- tight box of the rendering's alpha mask, square box x1.3, three ROIAlign crops, a batch-1 DINOv2 forward -- and
parks the normalised patch features on the HOST (`render_feats.append(....cpu())`, :231: 25 GB at the reference's 6000
views of ViT-B/14 tokens), from where the per-frame scoring (:295-296) reads them back.
Here, per batch of views: one `dh_roi_process_f32` call (boxes + mask / image / depth crops for the whole batch), one
batched DINOv2 forward (the model is the caller's: anything with `extract_features`, `smaller_edge_size`, `feat_size`
like ObjTracker/dino.py), and `dh_dino_prescale` writes the rows of the resident bf16 bank [N, P*D] that
`dino_match.dino_cos_topk` consumes -- the fp32 token tensor never leaves the device and is never kept.
Returned dict: the reference's keys (render_crop_imgs / render_crop_masks / render_crop_depths / render_roi_Ks /
render_feats_masks / render_rotations / render_translations / render_imgs) plus "templ_bank"; "render_feats"
(fp32 [N,P,D], normalised, on the device) only with keep_fp32=True.
"""
import ctypes

import torch
import torch.nn.functional as F

from . import _lib
from .camera import get_K_crop_resize
from .constants import REND_SIZE
from .preprocess import BBOX_EXPANSION, BBOX_PAD


def crop_views(renderings, depths=None, size=REND_SIZE):
    """renderings [n,H,W,4] f32 RGBA (alpha == 1 on the object, :198), depths [n,H,W,1] or [n,H,W] f32 or None ->
    dict(square_bbox [n,4] xywh, crop_mask [n,S,S] bool, crop_image [n,3,S,S] white outside the mask, crop_depth
    [n,1,S,S]) on the device (:198-215)."""
    if not torch.cuda.is_available():
        raise _lib.DynhorError("crop_views needs a CUDA device (dynhor_b200 has no CPU fallback)")
    r = renderings.cuda().float().contiguous()
    n, H, W, C = r.shape
    bits = (r[..., -1] == 1).to(torch.uint8).contiguous()
    d = None
    if depths is not None:
        d = depths.cuda().float().reshape(n, H, W).contiguous()
    dev, S = r.device, int(size)
    bounds = torch.empty(n, 4, dtype=torch.int32, device=dev)
    bbox = torch.empty(n, 4, device=dev)
    square = torch.empty(n, 4, device=dev)
    crop_mask = torch.empty(n, S, S, dtype=torch.uint8, device=dev)
    crop_image = torch.empty(n, 3, S, S, device=dev)
    crop_depth = torch.empty(n, S, S, device=dev) if d is not None else None
    _lib.check(_lib.load().dh_roi_process_f32(_lib.ptr(bits), _lib.ptr(r), C, _lib.ptr(d), n, H, W, S, BBOX_PAD,
                                              BBOX_EXPANSION, _lib.ptr(bounds), _lib.ptr(bbox), _lib.ptr(square),
                                              _lib.ptr(crop_mask), _lib.ptr(crop_image), _lib.ptr(crop_depth),
                                              _lib.stream_ptr()), "dh_roi_process_f32")
    empty = (bounds[:, 1] < 0).nonzero().flatten()
    if len(empty):   # torch.min of an empty tensor raises in the reference (:200)
        raise RuntimeError(f"min(): Expected reduction dim to be specified for input.numel() == 0 "
                           f"(template view {int(empty[0])} renders nothing)")
    out = {"bbox": bbox, "square_bbox": square, "crop_mask": crop_mask.view(torch.bool), "crop_image": crop_image}
    if crop_depth is not None:
        out["crop_depth"] = crop_depth[:, None]
    return out


def compute_prior_features(prior_infos, dino_model, batch_size=64, keep_fp32=False):
    """pose_initializtion.py:188-246, batched.  prior_infos: dict with "prior_batched_renderings" [N,H,W,4],
    "prior_depths" [N,H,W,1], "Ks" [N,3,3], "Rs", "Ts" (utils/render.py)."""
    lib = _lib.load()
    rend, depths, Ks = prior_infos["prior_batched_renderings"], prior_infos["prior_depths"], prior_infos["Ks"]
    N = len(rend)
    edge, fs = int(dino_model.smaller_edge_size), int(dino_model.feat_size)
    crops, masks, cdepths, roiK, fmasks, fp32 = [], [], [], [], [], []
    bank = None
    for a in range(0, N, batch_size):
        b = min(N, a + batch_size)
        c = crop_views(torch.as_tensor(rend[a:b]), torch.as_tensor(depths[a:b]))
        sq = c["square_bbox"].cpu()
        boxes = torch.stack([sq[:, 0], sq[:, 1], sq[:, 0] + sq[:, 2], sq[:, 1] + sq[:, 2]], 1)
        roiK.append(get_K_crop_resize(torch.as_tensor(Ks[a:b]).float().cpu(), boxes, [REND_SIZE]).cuda())   # :217-219
        with torch.no_grad():
            tokens = dino_model.extract_features(F.interpolate(c["crop_image"], edge, mode="bicubic",
                                                               align_corners=True))                              # :221-222
            fm = F.interpolate(c["crop_mask"][:, None].float(), fs, mode="nearest")[:, 0]                        # :224
        tokens = tokens.float().contiguous()
        n, P, D = tokens.shape
        if bank is None:
            bank = torch.empty(N, P * D, dtype=torch.bfloat16, device=tokens.device)
        # F.normalize(:223) and the division by the norms in the score (:296) collapse into one unit-vector scaling
        _lib.check(lib.dh_dino_prescale(_lib.ptr(tokens), None, n, P, D, ctypes.c_void_p(bank[a:b].data_ptr()),
                                        _lib.stream_ptr()), "dh_dino_prescale")
        if keep_fp32:
            fp32.append(F.normalize(tokens, dim=-1))
        crops.append(c["crop_image"].permute(0, 2, 3, 1))
        masks.append(c["crop_mask"])
        cdepths.append(c["crop_depth"])
        fmasks.append(fm)
    out = {
        "render_imgs": rend, "render_crop_imgs": torch.cat(crops), "render_crop_masks": torch.cat(masks),
        "render_crop_depths": torch.cat(cdepths), "render_roi_Ks": torch.cat(roiK),
        "render_feats_masks": torch.cat(fmasks), "render_rotations": prior_infos["Rs"],
        "render_translations": prior_infos["Ts"], "templ_bank": bank,
    }
    if keep_fp32:
        out["render_feats"] = torch.cat(fp32)
    return out
