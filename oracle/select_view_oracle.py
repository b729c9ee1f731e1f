"""oracle/select_view_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle; never a product path).

The candidate gating of ObjTracker/pose_initializtion.py:298-321, restated statement by statement for ONE frame
(the reference code is inline in find_optimal_pose, which cannot be imported here: it needs pytorch3d / detectron2 /
neural_renderer).  Inputs are what the reference has at that point: dino_cos [N] (:295-297), render_rotations [N,3,3]
(render_view_infos["render_rotations"]), rotations_init [1,3,3] (the previous frame's optimised rotation) or None,
former_max_idx, use_former.  Returns (max_idx, branch): max_idx as the reference leaves it (-1 = keep the previous
rotation, :323-325), branch = which path of :298-321 produced it (so that the tests can prove every path is hit).
Parity: pinned to the reference by construction of the restatement only (no reference test or fixture exists).
"""
import torch


def rotation_angle_difference(R1, R2):
    """utils/camera.py:4-9."""
    R_rel = R1 @ R2.transpose(1, 2)
    cos_theta = torch.clamp(0.5 * (torch.vmap(torch.trace)(R_rel) - 1), -1.0, 1.0)
    return (180.0 / torch.pi) * (torch.acos(cos_theta))


def select_view(dino_cos, render_rotations, rotations_init=None, former_max_idx=None, use_former=True):
    if not use_former or rotations_init is None:                                                  # :298
        return int(torch.argmax(dino_cos)), "argmax"                                              # :299
    rel_angle_full = rotation_angle_difference(rotations_init.clone(),                            # :301
                                               render_rotations.transpose(1, 2).clone())
    if former_max_idx != -1:                                                                      # :302
        former_rel_angle_full = rotation_angle_difference(                                        # :303-304
            render_rotations[former_max_idx:former_max_idx + 1].transpose(1, 2).clone(),
            render_rotations.transpose(1, 2).clone())
        cos_topk_num = 5                                                                          # :305
    else:
        former_rel_angle_full = torch.zeros_like(rel_angle_full)                                  # :307
        cos_topk_num = 10                                                                         # :308
    _, indices = torch.topk(dino_cos, cos_topk_num, largest=True)                                 # :309
    rel_angle = rel_angle_full[indices]                                                           # :310
    max_idx = indices[torch.argmin(rel_angle)].item()                                             # :311
    branch = "top5" if cos_topk_num == 5 else "top10"
    if rel_angle_full[max_idx] > 85.0 or former_rel_angle_full[max_idx] > 85.0:                   # :312
        branch += "+far_prev" if rel_angle_full[max_idx] > 85.0 else "+far_former"
        max_idx = -1                                                                              # :313
    if max_idx != -1:                                                                             # :314
        return int(max_idx), branch
    if torch.min(rel_angle_full) < 15.0:                                                          # :318
        max_idx = torch.argmin(rel_angle_full)                                                    # :319
        if (former_max_idx != -1 and former_rel_angle_full[max_idx].item() > 30.0) or \
                dino_cos[max_idx] < (torch.max(dino_cos) - torch.std(dino_cos)):                  # :320
            branch += "+near_rejected_former" if (former_max_idx != -1 and
                                                  former_rel_angle_full[max_idx].item() > 30.0) else "+near_rejected_cos"
            return -1, branch                                                                     # :321
        return int(max_idx), branch + "+near_accepted"
    return -1, branch + "+none_near"
