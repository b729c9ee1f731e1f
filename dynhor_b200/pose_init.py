"""Stage-1 silhouette term of ObjTracker's per-frame pose initialisation on the CUDA renderer.

Mirrors the silhouette-only part of `ObjTracker` in ObjTracker/pose_initializtion.py:
    __init__                 :36-110  (ref/keep masks, rotation/translation parameters, the anti_aliasing=False
                                       renderer at :98-105)
    apply_transformation     :112-117
    compute_offscreen_loss   :119-141
    coarse_forward           :143-155  (1 - IoU + 100000 * off-screen penalty)
and the optimisation loop of find_optimal_pose for mode="coarse" (:346-360).  The textured SoftPhong render ->
DINO -> semantic cosine term of `forward` (:157-186) is out of scope.
"""
import numpy as np
import torch
import torch.nn as nn

from .geometry import rot6d_to_matrix
from .losses import batch_mask_iou
from .renderer import Renderer, projection


class ObjTracker(nn.Module):
    """Silhouette-only ObjTracker: same constructor argument meaning as pose_initializtion.py:36-52 for the
    arguments the coarse path uses."""

    def __init__(self, ref_image, vertices, faces, rotation_init, translation_init, num_initializations=1, K=None):
        assert ref_image.shape[0] == ref_image.shape[1], "Must be square."
        super().__init__()
        self.register_buffer("vertices", vertices)
        self.register_buffer("faces", faces.repeat(num_initializations, 1, 1))
        ref_mask = torch.from_numpy((ref_image > 0).astype(np.float32))
        keep_mask = torch.from_numpy((ref_image >= 0).astype(np.float32))
        self.register_buffer("ref_mask", ref_mask.repeat(num_initializations, 1, 1))
        self.register_buffer("keep_mask", keep_mask.repeat(num_initializations, 1, 1))
        self.rotations = nn.Parameter(rotation_init.clone().float(), requires_grad=True)
        if rotation_init.shape[0] != translation_init.shape[0]:
            translation_init = translation_init.repeat(num_initializations, 1, 1)
        self.translations = nn.Parameter(translation_init.clone().float(), requires_grad=True)
        self.cuda()
        K = K.cuda()
        self.sil_renderer = Renderer(image_size=ref_image.shape[0], K=K, R=torch.eye(3).unsqueeze(0).cuda(),
                                     t=torch.zeros(1, 3).cuda(), orig_size=1, anti_aliasing=False)
        self.K = K

    def apply_transformation(self):
        rots = rot6d_to_matrix(self.rotations)
        return torch.matmul(self.vertices.repeat(rots.shape[0], 1, 1), rots) + self.translations

    def compute_offscreen_loss(self, verts):
        proj = projection(verts, self.sil_renderer.K, self.sil_renderer.R, self.sil_renderer.t,
                          self.sil_renderer.dist_coeffs, orig_size=1)
        coord_xy, coord_z = proj[:, :, :2], proj[:, :, 2:]
        zeros = torch.zeros_like(coord_z)
        lower_right = torch.max(coord_xy - 1, zeros).sum(dim=(1, 2))
        upper_left = torch.max(-1 - coord_xy, zeros).sum(dim=(1, 2))
        behind = torch.max(-coord_z, zeros).sum(dim=(1, 2))
        too_far = torch.max(coord_z - self.sil_renderer.far, zeros).sum(dim=(1, 2))
        return lower_right + upper_left + behind + too_far

    def coarse_forward(self):
        loss_dict = {}
        verts = self.apply_transformation()
        render_sil = self.sil_renderer(verts, self.faces, mode="silhouettes")
        render_mask = self.keep_mask * render_sil
        loss_dict["iou"] = (1 - batch_mask_iou(render_mask, self.ref_mask))
        with torch.no_grad():
            iou = batch_mask_iou(render_mask.detach(), self.ref_mask.detach())
        loss_dict["offscreen"] = 100000 * self.compute_offscreen_loss(verts)
        return loss_dict, iou


def optimize_coarse(model, num_iterations=50, lr=1e-3):
    """The loop of find_optimal_pose for mode="coarse" (pose_initializtion.py:346-358)."""
    optimizer = torch.optim.Adam(model.parameters(), lr=lr)
    history = []
    for _ in range(num_iterations):
        optimizer.zero_grad()
        loss_dict, iou = model.coarse_forward()
        losses = sum(loss_dict.values())
        loss = losses.sum()
        loss.backward()
        optimizer.step()
        history.append((float(loss.detach()), iou.detach().cpu().numpy().copy()))
    return history
