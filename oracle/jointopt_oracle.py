"""oracle/jointopt_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle; never a product path).

CPU restatement (torch CPU autograd + oracle/nmr_oracle.c) of the reference's joint pose optimisation:
  ObjTracker/jointopt.py:15-161          Joint_Optimizer / joint_optimize (state, forward, param groups, loop)
  ObjTracker/utils/losses.py:7-24,66-84  batch_mask_iou, compute_sil_loss, compute_smooth_loss
  ObjTracker/utils/geometry.py:7-38      rot6d_to_matrix / matrix_to_rot6d
  ObjTracker/utils/camera.py:179-207     compute_transformation_persp
  ObjTracker/utils/camera.py:26-63       projection (through oracle/nr_oracle.py)
Pinned against the reference's own Python run in the build container (tests/golden/make_golden.py imports
/root/reference/ObjTracker/{jointopt,utils/losses,utils/geometry,utils/camera}.py with only the missing
third-party rasteriser substituted by oracle/nr_oracle.py) -> tests/golden/jointopt_*.npz.
The rasteriser inside stays PARITY UNPINNED (see oracle/nmr_oracle.c).
"""
import time

import numpy as np
import torch
import torch.nn.functional as F

from . import nr_oracle

REND_SIZE = 256  # ObjTracker/utils/constants.py:2


def rot6d_to_matrix(rot_6d):
    """geometry.py:19-25. Gram-Schmidt; b1,b2,b3 become the COLUMNS of R. (cross over the last dim: the
    reference's dim-less torch.cross picks the first size-3 dim, identical unless B == 3.)"""
    rot_6d = rot_6d.view(-1, 3, 2)
    a1, a2 = rot_6d[:, :, 0], rot_6d[:, :, 1]
    b1 = F.normalize(a1)
    b2 = F.normalize(a2 - torch.einsum("bi,bi->b", b1, a2).unsqueeze(-1) * b1)
    b3 = torch.linalg.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), dim=-1)


def matrix_to_rot6d(rotmat):
    """geometry.py:38."""
    return rotmat.view(-1, 3, 3)[:, :, :2]


def transform_verts(verts_og, translations, rotations, scale_abs):
    """camera.py:195-207: (|s| * v) @ R + T, row-vector convention."""
    B = translations.shape[0]
    meshes = verts_og.repeat(B, 1, 1) if verts_og.ndimension() == 2 else verts_og
    return torch.matmul(scale_abs.view(-1, 1, 1) * meshes, rotations) + translations


def batch_mask_iou(ref, pred, eps=0.000001):
    """losses.py:7-24."""
    ref, pred = ref.float(), pred.float()
    if ref.max() > 1 or ref.min() < 0:
        raise ValueError("Ref mask should have values in [0, 1]")
    if pred.max() > 1 or pred.min() < 0:
        raise ValueError("Ref mask should have values in [0, 1]")
    inter = ref * pred
    union = ref + pred - inter
    return inter.sum(1).sum(1).float() / (union.sum(1).sum(1).float() + eps)


class JointOptOracle:
    """Same state and update rule as jointopt.py:15-62,125-141 on CPU."""

    def __init__(self, rot6d, trans, verts_og, faces, K_roi, target_masks, lr=1e-4, image_size=REND_SIZE,
                 anti_aliasing=True, optimize_object_scale=False, int_scale_init=1.0, correspondences=None,
                 corr_delta=1.0):
        # correspondences / corr_delta: the builder-defined reprojection term (oracle/corr_oracle.py); absent in
        # the reference, off unless given together with loss_weights["lw_corr_obj"] > 0
        self.correspondences = None if correspondences is None else torch.as_tensor(correspondences).float()
        self.corr_delta, self.image_size = float(corr_delta), int(image_size)
        self.translations_object = torch.nn.Parameter(torch.as_tensor(trans).float().reshape(-1, 1, 3).clone())
        self.rotations_object = torch.nn.Parameter(torch.as_tensor(rot6d).float().reshape(-1, 3, 2).clone())
        self.B = self.translations_object.shape[0]
        self.verts_object_og = torch.as_tensor(verts_og).float()
        faces = torch.as_tensor(faces).long()
        self.faces_object = faces if faces.ndim == 3 else faces[None].repeat(self.B, 1, 1)
        m = torch.as_tensor(target_masks).float()
        self.ref_mask_object = (m > 0).float()
        self.keep_mask_object = (m >= 0).float()
        self.camintr_rois_object = torch.as_tensor(K_roi).float().reshape(-1, 3, 3)
        scale = int_scale_init * torch.ones(1)
        self.optimize_object_scale = optimize_object_scale
        self.int_scales_object = torch.nn.Parameter(scale) if optimize_object_scale else scale
        self.renderer = nr_oracle.Renderer(image_size=image_size, K=self.camintr_rois_object,
                                           R=torch.eye(3).unsqueeze(0), t=torch.zeros(1, 3), orig_size=1,
                                           anti_aliasing=anti_aliasing)
        rigid = [self.translations_object] + ([self.int_scales_object] if optimize_object_scale else [])
        self.optimizer = torch.optim.Adam([{"params": rigid, "lr": lr},
                                           {"params": [self.rotations_object], "lr": lr * 10}])

    def get_verts_object(self):
        R = rot6d_to_matrix(self.rotations_object)
        return transform_verts(self.verts_object_og, self.translations_object, R, self.int_scales_object.abs())

    def render(self, verts=None):
        verts = self.get_verts_object() if verts is None else verts
        return self.renderer(verts, self.faces_object, mode="silhouettes")

    def forward(self, loss_weights=None):
        loss_dict, metric_dict = {}, {}
        verts = self.get_verts_object()
        if loss_weights is None or loss_weights["lw_smooth_obj"] > 0:
            loss_dict["loss_smooth_obj"] = ((verts[1:] - verts[:-1]) ** 2).mean()
        if loss_weights is None or loss_weights["lw_sil_obj"] > 0:
            rend = self.renderer(verts, self.faces_object, mode="silhouettes")
            image = self.keep_mask_object * rend
            l_m = torch.sum((image - self.ref_mask_object) ** 2) / self.keep_mask_object.sum()
            loss_dict["loss_sil_obj"] = (torch.zeros(1) + l_m) / len(verts)
            metric_dict["iou_object"] = batch_mask_iou(image, self.ref_mask_object).mean().item()
        if self.correspondences is not None and (loss_weights is None or loss_weights.get("lw_corr_obj", 0) > 0):
            from . import corr_oracle
            loss_dict["loss_corr_obj"] = corr_oracle.corr_loss(
                self.correspondences, rot6d_to_matrix(self.rotations_object), self.translations_object,
                self.int_scales_object.abs(), self.camintr_rois_object, self.image_size, self.corr_delta)
        return loss_dict, metric_dict

    def loss_and_grads(self, loss_weights):
        """One forward+backward without an optimiser step -> (dict of python floats, grad_rot6d, grad_trans)."""
        self.optimizer.zero_grad()
        loss_dict, metric_dict = self.forward(loss_weights)
        loss = sum(loss_dict[k] * loss_weights[k.replace("loss", "lw")] for k in loss_dict)
        loss.backward()
        out = {k: float(v.detach()) for k, v in loss_dict.items()}
        out.update(metric_dict)
        out["loss"] = float(loss.detach())
        g = {"rot6d": self.rotations_object.grad.detach().clone().numpy(),
             "trans": self.translations_object.grad.detach().clone().numpy()}
        if self.optimize_object_scale:
            g["scale"] = self.int_scales_object.grad.detach().clone().numpy()
        return out, g

    def step(self, loss_weights):
        out, g = self.loss_and_grads(loss_weights)
        self.optimizer.step()
        return out, g

    def run(self, loss_weights, num_iterations):
        """jointopt.py:142-161 loop -> loss_evolution dict of lists."""
        evo = {}
        for _ in range(num_iterations):
            out, _ = self.step(loss_weights)
            for k, v in out.items():
                evo.setdefault(k, []).append(v)
        return evo


def time_frame_iters(oracle, loss_weights, iters=1):
    t0 = time.perf_counter()
    for _ in range(iters):
        oracle.step(loss_weights)
    dt = time.perf_counter() - t0
    return oracle.B * iters / dt
