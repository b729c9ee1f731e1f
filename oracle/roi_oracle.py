"""CPU oracle of the ROI preprocessing that builds the joint optimisation's target masks (TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the product path).

Restates ObjTracker/run.py:26-72 (`process_input`) with its helpers
    utils/bbox.py:8-36      crop_and_resize            (ROIAlign((S, S), 1.0, 0, aligned=True) on N x C x H x W)
    utils/bbox.py:73-105    make_bbox_square, bbox_xy_to_wh / bbox_wh_to_xy (detectron2 BoxMode XYXY <-> XYWH)
    utils/maskutils.py:8-30 add_occlusions             (1 object, 0 background, -1 occluder that does not cover it)
and the two detectron2 pieces they call (`detectron2@v0.4`, requirements.txt:7, NOT installed and not under
/root/reference):
    detectron2/structures/masks.py  BitMasks.crop_and_resize = ROIAlign((S, S), 1.0, 0, aligned=True) on the bit
                                    masks as float32, then `>= 0.5`
    detectron2/layers/roi_align.py  ROIAlign.forward = torchvision.ops.roi_align(input, rois, output_size,
                                    spatial_scale, sampling_ratio, aligned)
torchvision IS in this image (0.26; the reference pins torch 2.1.0 -> torchvision 0.16, same CPU kernel), so the
ROIAlign arithmetic is not restated but executed: parity at this boundary is pinned to the library the reference
itself ends up calling.  Everything else is the reference's numpy / torch scalar code, line by line.
"""
import numpy as np
import torch
from torchvision.ops import roi_align

REND_SIZE = 256  # utils/constants.py:2


def bbox_xy_to_wh(b):  # utils/bbox.py:92-105 -> BoxMode.convert XYXY_ABS -> XYWH_ABS
    b = torch.as_tensor(b).clone()
    b[..., 2] -= b[..., 0]
    b[..., 3] -= b[..., 1]
    return b


def bbox_wh_to_xy(b):  # utils/bbox.py:108-117 -> BoxMode.convert XYWH_ABS -> XYXY_ABS
    if isinstance(b, np.ndarray):
        b = b.copy()
    else:
        b = torch.as_tensor(b).clone()
    b[..., 2] += b[..., 0]
    b[..., 3] += b[..., 1]
    return b


def make_bbox_square(bbox, bbox_expansion=0.0):  # utils/bbox.py:73-89 (numpy ops on the float32 tensor's values)
    bbox = np.asarray(torch.as_tensor(bbox))
    shape = bbox.shape
    bbox = bbox.reshape(-1, 4)
    center = np.stack((bbox[:, 0] + bbox[:, 2] / 2, bbox[:, 1] + bbox[:, 3] / 2), axis=1)
    b = np.expand_dims(np.maximum(bbox[:, 2], bbox[:, 3]), 1)
    b = b * np.float32(1 + bbox_expansion) if b.dtype == np.float32 else b * (1 + bbox_expansion)
    return np.hstack((center - b / 2, b, b)).reshape(shape)


def crop_and_resize(input_tensor, boxes, mask_size):  # utils/bbox.py:8-36
    batch_inds = torch.arange(len(boxes)).to(dtype=boxes.dtype)[:, None]
    rois = torch.cat([batch_inds, boxes], dim=1)
    return roi_align(input_tensor, rois.to(dtype=input_tensor.dtype), (mask_size, mask_size), 1.0, 0, True)


def bitmasks_crop_and_resize(bit_masks, boxes, mask_size):  # detectron2 BitMasks.crop_and_resize
    out = crop_and_resize(bit_masks.to(torch.float32)[:, None], boxes, mask_size).squeeze(1)
    return out >= 0.5


def add_occlusions(mask, occluder_mask, mask_bbox):  # utils/maskutils.py:8-30, one object
    bbox_mask = bbox_wh_to_xy(torch.Tensor(mask_bbox).unsqueeze(0))
    occlusions = bitmasks_crop_and_resize(occluder_mask, bbox_mask.repeat(occluder_mask.shape[0], 1), REND_SIZE)
    with_occlusions = torch.from_numpy(mask).float()
    with_occlusions[occlusions.sum(0) > 0] = -1
    with_occlusions[torch.from_numpy(mask)] = 1
    return with_occlusions.numpy()


def process_input(images, obj_masks, hand_masks):  # run.py:26-72
    objs = []
    for i, (obj_mask, hand_mask) in enumerate(zip(obj_masks, hand_masks)):
        obj_mask = (obj_mask == 255)
        hand_mask = (hand_mask == 255)
        hand_occlusions = torch.from_numpy(hand_mask).unsqueeze(0)
        bit_masks = torch.from_numpy(obj_mask).unsqueeze(0)
        nz = np.nonzero(obj_mask)
        min_row = max(np.min(nz[0]) - 5., 0)
        max_row = min(np.max(nz[0]) + 5., obj_mask.shape[0])
        min_col = max(np.min(nz[1]) - 5., 0)
        max_col = min(np.max(nz[1]) + 5., obj_mask.shape[1])
        box = torch.tensor([min_col, min_row, max_col, max_row]).float()
        bbox = bbox_xy_to_wh(box)
        square_bbox = make_bbox_square(bbox, 0.3)
        square_boxes = torch.FloatTensor(np.tile(bbox_wh_to_xy(square_bbox), (1, 1)))
        crop_masks = bitmasks_crop_and_resize(bit_masks, square_boxes, REND_SIZE)[0]
        obj = {"bbox": bbox, "class_id": -1, "score": None, "square_bbox": square_bbox,
               "crop_mask": crop_masks.numpy()}
        if images is not None:
            img = torch.from_numpy((images[i] / 255.).astype(np.float32)).permute(2, 0, 1).unsqueeze(0)
            images_crop = crop_and_resize(img, square_boxes, REND_SIZE)[0].permute(1, 2, 0)
            images_crop[crop_masks == 0] = torch.ones(3)
            obj["crop_image"] = images_crop.permute(2, 0, 1).numpy()
        obj["target_crop_mask"] = add_occlusions(obj["crop_mask"], hand_occlusions, obj["square_bbox"])
        objs.append(obj)
    return objs
