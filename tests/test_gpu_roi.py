"""GPU: dh_roi_process (dynhor_b200.preprocess, through the C ABI) against the oracle of run.py:26-72, which executes
torchvision's CPU roi_align (what detectron2's ROIAlign calls).  Everything bit-exact: boxes, crop masks, target
masks, image crops."""
import numpy as np
import pytest
import torch

from helpers import roi_scenes
from oracle import roi_oracle as ro

pytestmark = pytest.mark.gpu


def _check(images, objs, hands):
    from dynhor_b200.preprocess import process_input
    ref = ro.process_input(images, objs, hands)
    out = process_input(images, objs, hands)
    assert len(out) == len(ref)
    for b, (o, r) in enumerate(zip(out, ref)):
        assert set(o.keys()) == set(r.keys())
        assert o["class_id"] == -1 and o["score"] is None
        assert torch.equal(o["bbox"], r["bbox"]), b
        assert np.array_equal(o["square_bbox"], r["square_bbox"]) and o["square_bbox"].dtype == np.float32, b
        assert np.array_equal(o["crop_mask"], r["crop_mask"]) and o["crop_mask"].dtype == bool, b
        assert np.array_equal(o["target_crop_mask"], r["target_crop_mask"]), b
        assert np.array_equal(o["crop_image"], r["crop_image"]), b


@pytest.mark.parametrize("H,W,seed,border", [(480, 640, 0, False), (480, 640, 1, True), (270, 333, 2, True),
                                             (1080, 1920, 3, False)])
def test_process_input_bit_exact_vs_oracle(H, W, seed, border):
    """480x640 and 1080x1920 (BASELINE configs), objects against the image border, a width that is not a multiple of
    16 (scalar bounds path)."""
    _check(*roi_scenes(3, H, W, seed, border))


def test_batched_outputs_feed_joint_optimize_formats():
    from dynhor_b200.preprocess import process_input_batched
    images, objs, hands = roi_scenes(5, 480, 640, 7)
    r = process_input_batched(images, objs, hands)
    assert r["target_crop_mask"].shape == (5, 256, 256) and r["target_crop_mask"].is_cuda
    assert torch.equal(r["target_tri"].float(), r["target_crop_mask"])
    assert torch.equal(r["crop_mask"], r["target_crop_mask"] > 0)
    m = process_input_batched(None, objs, None)           # masks only, no occluder
    assert "crop_image" not in m and torch.equal(m["crop_mask"], r["crop_mask"]) and (m["target_tri"] >= 0).all()
    # device tensors in, same result
    ob = torch.from_numpy(np.stack(objs)).cuda()
    hb = torch.from_numpy(np.stack(hands)).cuda()
    r2 = process_input_batched(torch.from_numpy(np.stack(images)).cuda(), ob, hb)
    for k in r:
        assert torch.equal(r[k], r2[k]), k


def test_empty_object_mask_raises_like_the_reference():
    from dynhor_b200.preprocess import process_input
    images, objs, hands = roi_scenes(2, 120, 160, 9)
    objs[1] = np.zeros_like(objs[1])
    with pytest.raises(ValueError):
        ro.process_input(images, objs, hands)
    with pytest.raises(ValueError):
        process_input(images, objs, hands)
