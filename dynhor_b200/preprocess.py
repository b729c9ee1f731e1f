"""ROI preprocessing of ObjTracker/run.py on the GPU: the target masks (and crops) the joint optimisation consumes.

Host-side mirror of
    run.py:26-72               process_input(images, obj_masks, hand_masks) -> list of per-frame dicts
    utils/bbox.py:8-36,73-117  crop_and_resize / make_bbox_square / box modes
    utils/maskutils.py:8-30    add_occlusions
over `dh_roi_process` (csrc/dh_roi.cu): one call for all frames instead of a Python loop with three ROIAlign calls
per frame.  Same keys, shapes, dtypes and error behaviour as the reference (an empty object mask raises ValueError
like np.min does there).  `process_input_batched` keeps everything on the device for `joint_optimize`.
No CPU fallback.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .constants import REND_SIZE

BBOX_PAD = 5.0          # run.py:37-40
BBOX_EXPANSION = 0.3    # run.py:43


def _bits(masks, device, are_bits=False):
    """Masks with the SAM convention (255 = set, run.py:30-31) -> [B,H,W] uint8 0/1 on the device.  Host arrays
    (run.py's load_data yields float64 H x W arrays) are compared frame by frame straight into one pinned uint8
    buffer, so that 1 byte per pixel crosses PCIe instead of 8.  `are_bits`: the input already holds 0/1."""
    if isinstance(masks, torch.Tensor):
        m = masks.to(device, non_blocking=True)
        if are_bits and m.dtype in (torch.uint8, torch.bool):
            return m.contiguous().view(torch.uint8)
        return (m != 0 if are_bits else m == 255).to(torch.uint8).contiguous()
    first = np.asarray(masks[0])
    buf = torch.empty((len(masks),) + first.shape, dtype=torch.uint8).pin_memory()
    view = buf.numpy().view(np.bool_)
    for i, x in enumerate(masks):
        if are_bits:
            np.not_equal(np.asarray(x), 0, out=view[i])
        else:
            np.equal(np.asarray(x), 255, out=view[i])
    return buf.to(device, non_blocking=True)


def _images(images, device):
    if isinstance(images, torch.Tensor):
        return images.to(device, non_blocking=True).to(torch.uint8).contiguous()
    first = np.asarray(images[0])
    buf = torch.empty((len(images),) + first.shape, dtype=torch.uint8).pin_memory()
    view = buf.numpy()
    for i, x in enumerate(images):
        view[i] = x
    return buf.to(device, non_blocking=True)


def process_input_batched(images, obj_masks, hand_masks, size=REND_SIZE, device="cuda", with_images=True,
                          masks_are_bits=False):
    """All frames in one call.  Returns a dict of device tensors:
    bbox [B,4] xywh, square_bbox [B,4] xywh, crop_mask [B,S,S] bool, target_crop_mask [B,S,S] f32 in {1,0,-1},
    target_tri [B,S,S] int8, crop_image [B,3,S,S] f32 (if images are given)."""
    if not torch.cuda.is_available():
        raise _lib.DynhorError("process_input needs a CUDA device (dynhor_b200 has no CPU fallback)")
    lib = _lib.load()
    dev = torch.device(device)
    ob = _bits(obj_masks, dev, masks_are_bits)
    hb = _bits(hand_masks, dev, masks_are_bits) if hand_masks is not None else None
    B, H, W = ob.shape
    if hb is not None and hb.shape != ob.shape:
        raise AssertionError("object and hand masks must have the same shape")
    im = None
    if images is not None and with_images:
        im = _images(images, dev)
        if im.shape != (B, H, W, 3):
            raise AssertionError(f"images must be [B,H,W,3], got {tuple(im.shape)}")
    S = int(size)
    bounds = torch.empty(B, 4, dtype=torch.int32, device=dev)
    bbox = torch.empty(B, 4, dtype=torch.float32, device=dev)
    square = torch.empty(B, 4, dtype=torch.float32, device=dev)
    crop_mask = torch.empty(B, S, S, dtype=torch.uint8, device=dev)
    target = torch.empty(B, S, S, dtype=torch.float32, device=dev)
    tri = torch.empty(B, S, S, dtype=torch.int8, device=dev)
    crop_image = torch.empty(B, 3, S, S, dtype=torch.float32, device=dev) if im is not None else None
    _lib.check(lib.dh_roi_process(_lib.ptr(ob), _lib.ptr(hb), _lib.ptr(im), B, H, W, S, BBOX_PAD, BBOX_EXPANSION,
                                  _lib.ptr(bounds), _lib.ptr(bbox), _lib.ptr(square), _lib.ptr(crop_mask),
                                  _lib.ptr(target), _lib.ptr(tri), _lib.ptr(crop_image), _lib.stream_ptr()),
               "dh_roi_process")
    empty = (bounds[:, 1] < 0).nonzero().flatten()
    if len(empty):   # np.min(non_zero_indices[0]) on an empty mask (run.py:36)
        raise ValueError(f"zero-size array to reduction operation minimum which has no identity "
                         f"(empty object mask in frame {int(empty[0])})")
    out = {"bbox": bbox, "square_bbox": square, "crop_mask": crop_mask.view(torch.bool), "target_crop_mask": target,
           "target_tri": tri}
    if crop_image is not None:
        out["crop_image"] = crop_image
    return out


def process_input(images, obj_masks, hand_masks):
    """run.py:26-72 -- same signature and per-frame dicts (numpy arrays, torch `bbox`), computed in one GPU call."""
    r = process_input_batched(images, obj_masks, hand_masks)
    bbox = r["bbox"].cpu()
    square = r["square_bbox"].cpu().numpy()
    crop_mask = r["crop_mask"].cpu().numpy()
    target = r["target_crop_mask"].cpu().numpy()
    crop_image = r["crop_image"].cpu().numpy() if "crop_image" in r else None
    objs = []
    for b in range(len(bbox)):
        obj = {"bbox": bbox[b], "class_id": -1, "score": None, "square_bbox": square[b], "crop_mask": crop_mask[b]}
        if crop_image is not None:
            obj["crop_image"] = crop_image[b]
        obj["target_crop_mask"] = target[b]
        objs.append(obj)
    return objs
