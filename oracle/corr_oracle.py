"""oracle/corr_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle; never a product path).

[BUILDER-DEFINED, PARITY UNPINNED]  The dense-correspondence reprojection term named by BASELINE.json's north_star
does not exist in the reference (no DKM / correspondence / reprojection code under
/root/reference; README.md:43 only lists a data folder of the unreleased reconstruction stage).  This file is the
normative definition the CUDA kernel (dynhor_b200/csrc/dh_corr.cu) is tested against; there is nothing in the
reference to pin it to.  It reuses the reference's conventions where they exist:
  rigid transform   (|s| X) R + T                      ObjTracker/utils/camera.py:204-206
  projection        x/(z+1e-9), [u v 1] = [x y 1] K^T  ObjTracker/utils/camera.py:39-57 (before the v flip), K_roi in
                    unit-image coordinates             ObjTracker/pose_initializtion.py:327
  weighting         loss_corr_obj * lw_corr_obj        ObjTracker/jointopt.py:147-150
Plain torch ops (CPU autograd gives the gradients).
"""
import torch


def corr_frame_sums(records, rotations, translations, scale_abs, K_roi, image_size=256, delta=1.0):
    """records [B,C,6] (X[3], target u,v in ROI unit-image coordinates, weight w); rotations [B,3,3] (row-vector
    convention, columns b1 b2 b3), translations [B,1,3], scale_abs [1], K_roi [B,3,3] -> [B] sums of w * huber."""
    X, t, w = records[..., 0:3], records[..., 3:5], records[..., 5]
    c = torch.matmul(scale_abs.view(-1, 1, 1) * X, rotations) + translations.reshape(-1, 1, 3)
    zc = c[..., 2] + 1e-9
    x_, y_ = c[..., 0] / zc, c[..., 1] / zc
    K = K_roi.reshape(-1, 1, 3, 3)
    u = K[..., 0, 0] * x_ + K[..., 0, 1] * y_ + K[..., 0, 2]
    v = K[..., 1, 0] * x_ + K[..., 1, 1] * y_ + K[..., 1, 2]
    e = float(image_size) * torch.stack([u - t[..., 0], v - t[..., 1]], -1)
    r2 = (e * e).sum(-1)
    quad = r2 <= delta * delta
    r = torch.sqrt(torch.where(quad, torch.ones_like(r2), r2))      # sqrt only where it is differentiable
    rho = torch.where(quad, 0.5 * r2, delta * (r - 0.5 * delta))
    return (w * rho).sum(-1)


def corr_loss(records, rotations, translations, scale_abs, K_roi, image_size=256, delta=1.0, w_sum=None):
    """loss_corr_obj = sum_{b,c} w huber_delta(|e|) / sum_{b,c} w."""
    s = corr_frame_sums(records, rotations, translations, scale_abs, K_roi, image_size, delta).sum()
    w_sum = records[..., 5].double().sum().item() if w_sum is None else w_sum
    return s / w_sum
