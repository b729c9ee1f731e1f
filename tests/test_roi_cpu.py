"""CPU: the ROI preprocessing arithmetic the CUDA kernels run (dynhor_b200/csrc/dh_roi_core.h, built for the host by
tests/emu) against the oracle, which executes torchvision's roi_align -- the routine detectron2's ROIAlign, and with
it the reference's run.py:26-72, ends up calling.  Boxes, crop masks, target masks and image crops: bit-exact."""
import numpy as np
import pytest

import emu_lib as E
from helpers import roi_scenes
from oracle import roi_oracle as ro


def _compare(images, objs, hands, S=256):
    ref = ro.process_input(images, objs, hands)
    ob = np.stack([o == 255 for o in objs])
    hb = np.stack([h == 255 for h in hands])
    out = E.roi_process(ob, hb, np.stack(images) if images is not None else None, S)
    assert not out["status"].any()
    for b, r in enumerate(ref):
        assert np.array_equal(out["bbox"][b], r["bbox"].numpy()), b
        assert np.array_equal(out["square_bbox"][b], r["square_bbox"]), b
        assert np.array_equal(out["crop_mask"][b], r["crop_mask"]), b
        assert np.array_equal(out["target_crop_mask"][b], r["target_crop_mask"]), b
        if images is not None:
            assert np.array_equal(out["crop_image"][b], r["crop_image"]), b
    return ref, out


@pytest.mark.parametrize("H,W,seed,border", [(480, 640, 0, False), (480, 640, 1, True), (270, 333, 2, True)])
def test_emu_roi_preprocessing_bit_exact_vs_oracle(H, W, seed, border):
    images, objs, hands = roi_scenes(4, H, W, seed, border)
    ref, out = _compare(images, objs, hands)
    t = np.stack([r["target_crop_mask"] for r in ref])
    assert set(np.unique(t)) <= {-1.0, 0.0, 1.0} and (t == 1).any() and (t == 0).any()


def test_oracle_output_structure_matches_run_py():
    """Keys, dtypes and shapes of run.py:52-70's per-frame dict."""
    images, objs, hands = roi_scenes(1, 240, 320, 3)
    o = ro.process_input(images, objs, hands)[0]
    assert o["class_id"] == -1 and o["score"] is None
    assert tuple(o["bbox"].shape) == (4,) and o["square_bbox"].shape == (4,) and o["square_bbox"].dtype == np.float32
    assert o["crop_mask"].shape == (256, 256) and o["crop_mask"].dtype == bool
    assert o["crop_image"].shape == (3, 256, 256) and o["crop_image"].dtype == np.float32
    assert o["target_crop_mask"].shape == (256, 256) and o["target_crop_mask"].dtype == np.float32
    # the occluder never overwrites the object, and the crop is white outside the object (run.py:51, maskutils.py:27)
    assert (o["target_crop_mask"][o["crop_mask"]] == 1).all()
    assert (o["crop_image"][:, ~o["crop_mask"]] == 1).all()
    x, y, b, b2 = o["square_bbox"]
    assert b == b2 and b >= 1.3 * max(float(o["bbox"][2]), float(o["bbox"][3])) - 1e-3


def test_emu_reports_empty_object_mask():
    ob = np.zeros((1, 64, 64), np.uint8)
    assert E.roi_process(ob, None, None, 32)["status"][0] == 1
