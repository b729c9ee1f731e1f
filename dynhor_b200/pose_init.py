"""Stage 1 of ObjTracker -- the silhouette term of the per-frame pose initialisation -- on the fused CUDA iteration
(SURVEY.md 8f rank 1).

Reference behaviour covered (pose_initializtion.py):
    ObjTracker(ref_image, vertices, faces, ..., rotation_init, translation_init, num_initializations, K)   :36-110
        .apply_transformation()        :112-117   vertices @ R(rot6d) + T
        .compute_offscreen_loss(verts) :119-141   per-candidate off-screen penalty
        .coarse_forward()              :143-155   ({"iou": 1 - IoU, "offscreen": 100000 * penalty}, IoU)
    the mode="coarse" loop of find_optimal_pose       :346-360   Adam(model.parameters(), lr), num_iterations steps,
                                                                 candidates sorted by their last loss (:368-372)
The textured SoftPhong render -> DINOv2 -> semantic cosine term of `forward` (:157-186) is not built (SURVEY.md 8f
rank 3: ViT forward/backward + a textured soft rasteriser).

Two ways in, one set of kernels:
  * `coarse_optimize(...)`: every (frame, candidate) pair of a whole sequence in ONE fused run -- anti_aliasing=False
    raster, IoU loss and its two-valued gradient, off-screen penalty on the projected vertices, edge-scan backward,
    single-group Adam, all inside dh_jointopt_run with loss_mode DH_LOSS_STAGE1, replayed as a CUDA graph.  The
    candidates are independent, so a sequence is initialised in one batch when their starting poses are known up
    front (use_former=False, or the poses of a previous pass).
  * `ObjTracker`: the reference's module for one frame; `coarse_forward()` is the composable autograd path (CUDA
    renderer op + torch glue), `optimize()` runs the same fused iteration on the module's own parameters.
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .geometry import rot6d_to_matrix
from .losses import batch_mask_iou, offscreen_penalty
from .renderer import Renderer, projection

OFFSCREEN_WEIGHT = 100000.0   # pose_initializtion.py:154


class _Candidates(nn.Module):
    """Parameter / buffer holder with the attribute names FusedJointOpt reads (those of Joint_Optimizer)."""

    def __init__(self, rot6d, trans, verts, faces, K, target_masks):
        super().__init__()
        if not torch.cuda.is_available():
            raise _lib.DynhorError("stage-1 optimisation needs a CUDA device (dynhor_b200 has no CPU fallback)")
        self.rotations_object = nn.Parameter(rot6d.detach().clone().float().reshape(-1, 3, 2).contiguous().cuda())
        self.translations_object = nn.Parameter(trans.detach().clone().float().reshape(-1, 1, 3).contiguous().cuda())
        m = target_masks.cuda()
        self.register_buffer("ref_mask_object", (m > 0).float())
        self.register_buffer("keep_mask_object", (m >= 0).float())
        self.register_buffer("verts_object_og", verts.float().cuda())
        self.register_buffer("faces_object", faces.cuda())
        self.register_buffer("camintr_rois_object", K.float().cuda())
        self.register_buffer("int_scales_object", torch.ones(1).cuda())
        self.optimize_object_scale = False
        self.corr_term = None


def coarse_optimize(target_masks, vertices, faces, rot6d_init, translation_init, K, num_iterations=100, lr=1e-2,
                    use_graph=True, return_history=True):
    """All candidates of a sequence at once.
        target_masks [B,S,S] in {1 object, 0 background, -1 occluder}; vertices [V,3]; faces [F,3];
        rot6d_init [B,3,2]; translation_init [B,1,3]; K [B,3,3] or [1,3,3] ROI intrinsics in unit-image coordinates
        (one row per candidate: repeat a frame's mask / K for several candidates of that frame).
    Returns dict(rotations [B,3,2], translations [B,1,3], losses [B] (each candidate's loss at the last iteration,
    what :368-372 sorts by), iou [B], history)."""
    from .jointopt import FusedJointOpt
    B = int(rot6d_init.shape[0])
    masks = torch.as_tensor(target_masks)
    K = torch.as_tensor(K).reshape(-1, 3, 3)
    model = _Candidates(torch.as_tensor(rot6d_init), torch.as_tensor(translation_init), torch.as_tensor(vertices),
                        torch.as_tensor(faces), K.expand(B, 3, 3).contiguous(), masks.expand(B, *masks.shape[-2:]))
    lw = {"lw_sil_obj": 1.0, "lw_offscreen": OFFSCREEN_WEIGHT}
    with FusedJointOpt(model, lw, lr, num_iterations, stage1=True) as fused:
        fused.run(num_iterations, use_graph=use_graph)
        last = fused.frame_losses()
        history = fused.history() if return_history else None
    return {"rotations": model.rotations_object.detach(), "translations": model.translations_object.detach(),
            "losses": last["loss"].float(), "iou": last["iou"].float(), "history": history}


class ObjTracker(nn.Module):
    """One frame's candidates, with the reference module's constructor keywords, parameter names (`rotations`
    [n,3,2], `translations` [n,1,3]) and method names.  Arguments of the textured path (textures, dino_model,
    gt_dino_feat, rasterizer, shader, lw_mask, lw_sem) are accepted and unused."""

    def __init__(self, ref_image, vertices, faces, textures=None, dino_model=None, gt_dino_feat=None,
                 rotation_init=None, translation_init=None, num_initializations=1, K=None, rasterizer=None,
                 shader=None, lw_mask=1.0, lw_sem=1.0):
        side = ref_image.shape[0]
        assert side == ref_image.shape[1], "Must be square."
        super().__init__()
        if not torch.cuda.is_available():
            raise _lib.DynhorError("ObjTracker needs a CUDA device (dynhor_b200 has no CPU fallback)")
        n = int(num_initializations)
        target = torch.as_tensor(np.asarray(ref_image), dtype=torch.float32)
        self.register_buffer("target", target)                                      # {-1, 0, 1}, [S,S]
        self.register_buffer("ref_mask", (target > 0).float().expand(n, side, side).contiguous())
        self.register_buffer("keep_mask", (target >= 0).float().expand(n, side, side).contiguous())
        self.register_buffer("vertices", torch.as_tensor(vertices).float())
        f = torch.as_tensor(faces)
        self.register_buffer("faces", (f if f.ndim == 3 else f[None]).expand(n, -1, 3).contiguous())
        rot = torch.as_tensor(rotation_init).float()
        tr = torch.as_tensor(translation_init).float()
        if tr.shape[0] != rot.shape[0]:
            tr = tr.expand(rot.shape[0], *tr.shape[1:])
        self.rotations = nn.Parameter(rot.clone())
        self.translations = nn.Parameter(tr.clone())
        self.cuda()
        self.K = torch.as_tensor(K).float().cuda()
        self.sil_renderer = Renderer(image_size=side, K=self.K, R=torch.eye(3, device="cuda")[None],
                                     t=torch.zeros(1, 3, device="cuda"), orig_size=1, anti_aliasing=False)
        self.best_score, self.former_max_idx, self.losses = 0, None, None

    def apply_transformation(self):
        return torch.matmul(self.vertices.expand(self.rotations.shape[0], -1, -1),
                            rot6d_to_matrix(self.rotations)) + self.translations

    def compute_offscreen_loss(self, verts):
        r = self.sil_renderer
        return offscreen_penalty(projection(verts, r.K, r.R, r.t, r.dist_coeffs, orig_size=1), r.far)

    def coarse_forward(self):
        verts = self.apply_transformation()
        seen = self.keep_mask * self.sil_renderer(verts, self.faces, mode="silhouettes")
        iou = batch_mask_iou(seen, self.ref_mask)
        terms = {"iou": 1 - iou, "offscreen": OFFSCREEN_WEIGHT * self.compute_offscreen_loss(verts)}
        return terms, iou.detach()

    def forward(self):
        raise NotImplementedError("the textured SoftPhong -> DINOv2 -> semantic loss path (pose_initializtion.py:"
                                  "157-186) is not built (SURVEY.md 8f rank 3); use coarse_forward() / optimize()")

    def optimize(self, num_iterations=100, lr=1e-2, sort_best=True, use_graph=True):
        """The mode="coarse" loop of find_optimal_pose (:346-372) as one fused run on this module's parameters:
        afterwards `rotations` / `translations` hold the candidates (best first when sort_best) and `losses` their
        losses at the last iteration.  Returns the loss history."""
        n = self.rotations.shape[0]
        out = coarse_optimize(self.target.expand(n, -1, -1), self.vertices, self.faces[0], self.rotations.detach(),
                              self.translations.detach(), self.K, num_iterations, lr, use_graph)
        order = torch.argsort(out["losses"]) if sort_best else torch.arange(n, device=out["losses"].device)
        with torch.no_grad():
            self.rotations.copy_(out["rotations"][order])
            self.translations.copy_(out["translations"][order])
        self.losses = out["losses"][order]
        return out["history"]


def optimize_coarse(model, num_iterations=50, lr=1e-3):
    """Composable path: the reference loop itself (:346-358) -- coarse_forward + loss.backward() + torch.optim.Adam --
    on the CUDA renderer op.  Returns [(total loss, per-candidate IoU)] per iteration."""
    optimizer = torch.optim.Adam(model.parameters(), lr=lr)
    trace = []
    for _ in range(num_iterations):
        optimizer.zero_grad()
        terms, iou = model.coarse_forward()
        total = sum(terms.values()).sum()
        total.backward()
        optimizer.step()
        trace.append((float(total.detach()), iou.cpu().numpy().copy()))
    return trace
