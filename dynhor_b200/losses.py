"""Loss terms of the composable (autograd) path, with the CUDA silhouette renderer underneath.

Interface mirror of ObjTracker/utils/losses.py -- same public names, argument meaning, return structure and error
behaviour, so that code written against the reference keeps working:

    batch_mask_iou(ref, pred, eps)        losses.py:7-24   per-frame IoU; ValueError when a mask leaves [0, 1]
    Losses(ref_mask_object, keep_mask_object, camintr_rois_object)             losses.py:26-40
      .compute_offscreen_loss(verts)      losses.py:42-64  summed distance of projected vertices outside the view
      .compute_sil_loss(verts, faces)     losses.py:66-78  ({"loss_sil_obj": Tensor[1]}, {"iou_object": float})
      .compute_smooth_loss(verts)         losses.py:80-84  {"loss_smooth_obj": Tensor[]}

`jointopt.joint_optimize` does not come through here: it runs the fused kernels (csrc/dh_jointopt.cu).
"""
import torch

from .constants import REND_SIZE
from .renderer import Renderer, projection


def _require_unit_range(mask):
    lo, hi = mask.min(), mask.max()
    if hi > 1 or lo < 0:
        raise ValueError("Ref mask should have values in [0, 1], " f"not [{lo, hi}]")


def batch_mask_iou(ref, pred, eps=0.000001):
    """Intersection over union of every frame's pair of soft masks, [B,H,W] x [B,H,W] -> [B]."""
    ref, pred = ref.float(), pred.float()
    _require_unit_range(ref)
    _require_unit_range(pred)
    both = ref * pred
    either = ref + pred - both
    per_frame = lambda t: t.sum(1).sum(1).float()   # rows first, then columns: the reference's summation order
    return per_frame(both) / (per_frame(either) + eps)


def offscreen_penalty(uvz, far):
    """[B,V,3] projected vertices (u, v in NDC, depth) -> [B]: by how much they leave [-1, 1]^2 x (0, far), summed over
    the vertices of each frame (losses.py:42-64 sums it over everything, pose_initializtion.py:119-141 per candidate)."""
    uv, z = uvz[..., :2], uvz[..., 2:]
    zero = torch.zeros_like(z)
    beyond = torch.max(uv - 1, zero).sum(dim=(1, 2)) + torch.max(-1 - uv, zero).sum(dim=(1, 2))
    depth = torch.max(-z, zero).sum(dim=(1, 2)) + torch.max(z - far, zero).sum(dim=(1, 2))
    return beyond + depth


class Losses():
    """Holds the target masks and the silhouette renderer of one sequence."""

    def __init__(self, ref_mask_object, keep_mask_object, camintr_rois_object, image_size=None):
        self.ref_mask_object, self.keep_mask_object = ref_mask_object, keep_mask_object
        self.camintr_rois_object = camintr_rois_object
        device = camintr_rois_object.device
        side = int(ref_mask_object.shape[-1]) if image_size is None else image_size
        self.sil_renderer = Renderer(image_size=side or REND_SIZE, K=camintr_rois_object,
                                     R=torch.eye(3, device=device)[None], t=torch.zeros(1, 3, device=device),
                                     orig_size=1)

    def compute_offscreen_loss(self, verts):
        """How far the projected vertices stick out of the view volume ([-1, 1]^2 x (0, far)), summed."""
        r = self.sil_renderer
        return offscreen_penalty(projection(verts, r.K, r.R, r.t, r.dist_coeffs, orig_size=1), r.far).sum()

    def compute_sil_loss(self, verts, faces):
        """Masked L2 between the rendered silhouettes and the target, per kept pixel and per frame, plus the IoU."""
        silhouettes = self.sil_renderer(verts, faces, mode="silhouettes")
        kept = self.keep_mask_object * silhouettes
        residual = torch.sum((kept - self.ref_mask_object) ** 2) / self.keep_mask_object.sum()
        loss = torch.zeros(1, device=verts.device, dtype=torch.float32) + residual
        iou = batch_mask_iou(kept, self.ref_mask_object).mean().item()
        return {"loss_sil_obj": loss / len(verts)}, {"iou_object": iou}

    def compute_smooth_loss(self, verts):
        """Mean squared vertex displacement between consecutive frames."""
        step = verts[1:] - verts[:-1]
        return {"loss_smooth_obj": (step ** 2).mean()}
