"""oracle/stage1_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle; never a product path).

CPU restatement (torch CPU autograd + oracle/nmr_oracle.c through oracle/nr_oracle.py) of the silhouette term of the
per-frame pose initialisation, ObjTracker/pose_initializtion.py:
    :36-110    ObjTracker.__init__ (ref / keep masks, rot6d + translation parameters, anti_aliasing=False renderer)
    :112-117   apply_transformation           :119-141  compute_offscreen_loss
    :143-155   coarse_forward  (1 - IoU, 100000 * off-screen penalty)
    :346-360   the mode="coarse" loop of find_optimal_pose (one Adam group, lr)
Pinned against the reference's own code: tests/golden/make_golden.py runs pose_initializtion.ObjTracker.coarse_forward
unmodified (with the oracle renderer in place of the absent neural_renderer) -> tests/golden/stage1_coarse.npz,
stage1_multi.npz; tests/test_oracle_golden.py checks this file against them.  The rasteriser inside stays PARITY
UNPINNED (oracle/nmr_oracle.c).
"""
import torch

from . import nr_oracle
from .jointopt_oracle import batch_mask_iou, rot6d_to_matrix

OFFSCREEN_WEIGHT = 100000.0


class Stage1Oracle:
    def __init__(self, target_masks, verts, faces, rot6d_init, trans_init, K, lr=1e-2):
        """target_masks [n,S,S] in {-1,0,1}; verts [V,3]; faces [F,3]; rot6d_init [n,3,2]; trans_init [n,1,3];
        K [n,3,3] unit-image ROI intrinsics."""
        t = lambda a, dt=torch.float32: torch.as_tensor(a).to(dt)  # noqa: E731
        m = t(target_masks)
        self.ref, self.keep = (m > 0).float(), (m >= 0).float()
        self.verts, self.faces = t(verts), t(faces, torch.int64)
        self.rotations = torch.nn.Parameter(t(rot6d_init).clone())
        self.translations = torch.nn.Parameter(t(trans_init).reshape(-1, 1, 3).clone())
        n, S = self.rotations.shape[0], m.shape[-1]
        self.K = t(K).reshape(-1, 3, 3).expand(n, 3, 3).contiguous()
        self.renderer = nr_oracle.Renderer(image_size=S, K=self.K, R=torch.eye(3)[None], t=torch.zeros(1, 3),
                                           orig_size=1, anti_aliasing=False)
        self.opt = torch.optim.Adam([self.rotations, self.translations], lr=lr)

    def losses(self):
        n = self.rotations.shape[0]
        verts = torch.matmul(self.verts.repeat(n, 1, 1), rot6d_to_matrix(self.rotations)) + self.translations   # :116
        sil = self.renderer(verts, self.faces[None].repeat(n, 1, 1), mode="silhouettes")                       # :146
        iou = batch_mask_iou(self.keep * sil, self.ref)                                                          # :148
        r = self.renderer
        proj = nr_oracle.projection(verts, r.K, r.R, r.t, r.dist_coeffs, orig_size=1)                            # :125
        xy, z = proj[:, :, :2], proj[:, :, 2:]
        zeros = torch.zeros_like(z)
        off = (torch.max(xy - 1, zeros).sum(dim=(1, 2)) + torch.max(-1 - xy, zeros).sum(dim=(1, 2))            # :134-141
               + torch.max(-z, zeros).sum(dim=(1, 2)) + torch.max(z - r.far, zeros).sum(dim=(1, 2)))
        return (1 - iou) + OFFSCREEN_WEIGHT * off, iou.detach(), off.detach()

    def step(self):
        """One iteration of :347-358.  Returns (per-candidate losses, IoU, off-screen penalty, gradients)."""
        self.opt.zero_grad()
        lv, iou, off = self.losses()
        lv.sum().backward()
        g = (self.rotations.grad.clone(), self.translations.grad.clone())
        self.opt.step()
        return lv.detach(), iou, off, g
