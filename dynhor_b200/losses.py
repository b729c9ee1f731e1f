"""Host-side mirror of ObjTracker/utils/losses.py with the CUDA renderer underneath.

Same class, method names, arguments, return structure and error behaviour:
    batch_mask_iou(ref, pred)                          losses.py:7-24   (ValueError outside [0,1])
    Losses(ref_mask_object, keep_mask_object, camintr_rois_object)      losses.py:26-40
    Losses.compute_sil_loss(verts, faces) -> ({"loss_sil_obj": Tensor[1]}, {"iou_object": float})   :66-78
    Losses.compute_smooth_loss(verts)     -> {"loss_smooth_obj": Tensor[]}                          :80-84
    Losses.compute_offscreen_loss(verts)                                                            :42-64
This is the composable (autograd) path; jointopt.joint_optimize uses the fused kernels instead.
"""
import torch

from .constants import REND_SIZE
from .renderer import Renderer, projection


def batch_mask_iou(ref, pred, eps=0.000001):
    ref = ref.float()
    pred = pred.float()
    if ref.max() > 1 or ref.min() < 0:
        raise ValueError("Ref mask should have values in [0, 1], " f"not [{ref.min(), ref.max()}]")
    if pred.max() > 1 or pred.min() < 0:
        raise ValueError("Ref mask should have values in [0, 1], " f"not [{pred.min(), pred.max()}]")
    inter = ref * pred
    union = ref + pred - inter
    ious = inter.sum(1).sum(1).float() / (union.sum(1).sum(1).float() + eps)
    return ious


class Losses():
    def __init__(self, ref_mask_object, keep_mask_object, camintr_rois_object, image_size=None):
        self.ref_mask_object = ref_mask_object
        self.keep_mask_object = keep_mask_object
        self.camintr_rois_object = camintr_rois_object
        dev = camintr_rois_object.device
        size = int(ref_mask_object.shape[-1]) if image_size is None else image_size
        self.sil_renderer = Renderer(image_size=size if size else REND_SIZE, K=camintr_rois_object,
                                     R=torch.eye(3, device=dev).unsqueeze(0), t=torch.zeros(1, 3, device=dev),
                                     orig_size=1)

    def compute_offscreen_loss(self, verts):
        proj = projection(verts, self.sil_renderer.K, self.sil_renderer.R, self.sil_renderer.t,
                          self.sil_renderer.dist_coeffs, orig_size=1)
        coord_xy, coord_z = proj[:, :, :2], proj[:, :, 2:]
        zeros = torch.zeros_like(coord_z)
        lower_right = torch.max(coord_xy - 1, zeros).sum()
        upper_left = torch.max(-1 - coord_xy, zeros).sum()
        behind = torch.max(-coord_z, zeros).sum()
        too_far = torch.max(coord_z - self.sil_renderer.far, zeros).sum()
        return lower_right + upper_left + behind + too_far

    def compute_sil_loss(self, verts, faces):
        loss_sil = torch.zeros(1, device=verts.device, dtype=torch.float32)
        rend = self.sil_renderer(verts, faces, mode="silhouettes")
        image = self.keep_mask_object * rend
        l_m = torch.sum((image - self.ref_mask_object) ** 2) / self.keep_mask_object.sum()
        loss_sil = loss_sil + l_m
        ious = batch_mask_iou(image, self.ref_mask_object)
        return {"loss_sil_obj": loss_sil / len(verts)}, {"iou_object": ious.mean().item()}

    def compute_smooth_loss(self, verts):
        smooth_loss_obj = ((verts[1:] - verts[:-1]) ** 2).mean()
        return {"loss_smooth_obj": smooth_loss_obj}
