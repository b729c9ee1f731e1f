"""oracle/prior_features_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle; never a product path).

compute_prior_features of ObjTracker/pose_initializtion.py:188-246 restated line by line on CPU: one template view at
a time, detectron2's BitMasks.crop_and_resize / ROIAlign executed through torchvision.ops.roi_align (oracle/
roi_oracle.py: the library the reference itself ends up calling), get_K_crop_resize restated from utils/camera.py:84-130.
`dino_model` is whatever the caller passes (the tests use a small deterministic stand-in for DINOv2: no network).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import roi_oracle as ro

REND_SIZE, BBOX_EXPANSION_FACTOR = 256, 0.3      # utils/constants.py:2-3


def get_K_crop_resize(K, boxes, crop_resize):     # utils/camera.py:84-130
    K = K.float()
    boxes = boxes.float()
    new_K = K.clone()
    final_width, final_height = max(crop_resize), min(crop_resize)
    crop_width = boxes[:, 2] - boxes[:, 0]
    crop_height = boxes[:, 3] - boxes[:, 1]
    crop_cj = (boxes[:, 0] + boxes[:, 2]) / 2
    crop_ci = (boxes[:, 1] + boxes[:, 3]) / 2
    cx = K[:, 0, 2] + (crop_width - 1) / 2 - crop_cj
    cy = K[:, 1, 2] + (crop_height - 1) / 2 - crop_ci
    center_x = (crop_width - 1) / 2
    center_y = (crop_height - 1) / 2
    orig_cx_diff = cx - center_x
    orig_cy_diff = cy - center_y
    scale_x = final_width / crop_width
    scale_y = final_height / crop_height
    scaled_center_x = (final_width - 1) / 2
    scaled_center_y = (final_height - 1) / 2
    fx = scale_x * K[:, 0, 0]
    fy = scale_y * K[:, 1, 1]
    cx = scaled_center_x + scale_x * orig_cx_diff
    cy = scaled_center_y + scale_y * orig_cy_diff
    new_K[:, 0, 0] = fx
    new_K[:, 1, 1] = fy
    new_K[:, 0, 2] = cx
    new_K[:, 1, 2] = cy
    return new_K


def compute_prior_features(prior_infos, dino_model):
    H, W = prior_infos["prior_batched_renderings"].shape[1:3]       # RENDER_H, RENDER_W (constants.py:4)
    imgs, masks, depths, feats, fmasks, Ks = [], [], [], [], [], []
    for rendering, depth, K_ in zip(prior_infos["prior_batched_renderings"], prior_infos["prior_depths"],
                                    prior_infos["Ks"]):
        render_mask = (rendering[:, :, -1] == 1)                                                    # :198
        nz = torch.nonzero(render_mask)                                                             # :200
        min_row = max(torch.min(nz[:, 0]) - 5., 0)                                                  # :201-204
        max_row = min(torch.max(nz[:, 0]) + 5., H)
        min_col = max(torch.min(nz[:, 1]) - 5., 0)
        max_col = min(torch.max(nz[:, 1]) + 5., W)
        box = torch.tensor([min_col, min_row, max_col, max_row]).float()
        bbox = ro.bbox_xy_to_wh(box)
        square_bbox = ro.make_bbox_square(bbox, BBOX_EXPANSION_FACTOR)
        square_boxes = torch.FloatTensor(np.tile(ro.bbox_wh_to_xy(square_bbox), (1, 1)))
        crop_mask = ro.bitmasks_crop_and_resize(render_mask.unsqueeze(0), square_boxes, REND_SIZE).clone()[0]   # :210
        crop_image = ro.crop_and_resize(rendering[:, :, :3].permute(2, 0, 1).unsqueeze(0), square_boxes,
                                        REND_SIZE).clone()[0].permute(1, 2, 0)                      # :212
        crop_depth = ro.crop_and_resize(depth.permute(2, 0, 1).unsqueeze(0), square_boxes, REND_SIZE).clone()[0]
        crop_image[crop_mask == 0] = torch.ones(3)                                                  # :215
        x, y, b, _ = square_bbox
        Ks.append(get_K_crop_resize(torch.Tensor(K_).unsqueeze(0), torch.tensor([[x, y, x + b, y + b]]), [REND_SIZE]))
        with torch.no_grad():
            f = dino_model.extract_features(F.interpolate(crop_image.permute(2, 0, 1).unsqueeze(0),
                                                          dino_model.smaller_edge_size, mode="bicubic",
                                                          align_corners=True))                      # :221
            f = F.normalize(f, dim=-1)                                                              # :223
            fm = F.interpolate(crop_mask[None, None].float(), dino_model.feat_size, mode="nearest")  # :224
        imgs.append(crop_image.unsqueeze(0)), masks.append(crop_mask.unsqueeze(0)), depths.append(crop_depth.unsqueeze(0))
        feats.append(f), fmasks.append(fm.squeeze(1))
    return {"render_crop_imgs": torch.cat(imgs), "render_crop_masks": torch.cat(masks),
            "render_crop_depths": torch.cat(depths), "render_roi_Ks": torch.cat(Ks),
            "render_feats": torch.cat(feats), "render_feats_masks": torch.cat(fmasks)}
