"""oracle/dino_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle; never a product path).

The view-selection score and top-k of ObjTracker/pose_initializtion.py:295-296,299,309 on CPU, expression kept
verbatim (one frame against N templates), fp32.  PARITY UNPINNED beyond that: the reference has no importable
function for it (it is inline code inside find_optimal_pose, which needs pytorch3d / detectron2 to import) and no
test or fixture; the expression itself is the contract.
"""
import torch


def dino_cos(gt_dino_feat, render_feats, cos_mask):
    """gt_dino_feat [1,P,D], render_feats [N,P,D], cos_mask [1,P] -> [N]   (pose_initializtion.py:295-296)."""
    return (cos_mask * torch.sum(gt_dino_feat * render_feats, dim=-1)
            / (torch.norm(gt_dino_feat, dim=-1) * torch.norm(render_feats, dim=-1) + 1e-6)).sum(1) / cos_mask.sum(1)


def dino_cos_topk(frame_feats, frame_masks, templ_feats, k):
    """All frames, sequentially like the reference loop.  -> scores [Fm,N], topk values, indices."""
    scores = torch.stack([dino_cos(frame_feats[f:f + 1], templ_feats, frame_masks[f:f + 1])
                          for f in range(frame_feats.shape[0])])
    vals, idx = torch.topk(scores, k, dim=1, largest=True)
    return scores, vals, idx
