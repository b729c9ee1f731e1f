"""Shared helpers of the test-suite."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
JOINT_CASES = ["s64_b5", "s256_b4", "s128_b6_lr", "s64_b4_scale"]


def load_golden(name):
    return np.load(os.path.join(GOLDEN, f"jointopt_{name}.npz"))


def unpack_alpha(abits):
    """[B,is,is/32] uint32 bitmaps -> [B,is,is] bool."""
    a = np.asarray(abits).view(np.uint32)
    return ((a[..., None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool).reshape(a.shape[0], a.shape[1], -1)


def golden_alpha(g):
    return np.unpackbits(g["orc_alpha0"], axis=-1).astype(bool)


def rot6d_to_rotmats(rot6d):
    """A [B,3,3] matrix whose first two columns are the 6D parameters (what matrix_to_rot6d slices back)."""
    a = np.asarray(rot6d, np.float32)
    c = np.cross(a[:, :, 0], a[:, :, 1])
    return np.concatenate([a, c[:, :, None]], -1).astype(np.float32)


def object_parameters_from_golden(g):
    import torch
    B = len(g["rot6d_init"])
    R = rot6d_to_rotmats(g["rot6d_init"])
    out = []
    for b in range(B):
        out.append({
            "rotations": torch.from_numpy(R[b:b + 1].copy()),
            "translations": torch.from_numpy(g["trans_init"][b:b + 1].copy()),
            "K_roi": torch.from_numpy(g["K_roi"][b:b + 1].copy()).unsqueeze(0),
            "target_masks": torch.from_numpy(g["target_masks"][b:b + 1].astype(np.float32)),
        })
    return out


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def roi_scenes(n, H=480, W=640, seed=0, border=False):
    """Synthetic SAM-style inputs of run.py:load_data: per frame an RGB image (uint8 HWC), an object mask and a hand
    mask (float64 arrays, 255 = set).  Rotated ellipses with a disc occluder; `border` pushes objects against the
    image edges (box clamping, samples outside the image)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    images, objs, hands = [], [], []
    for i in range(n):
        if border and i % 2 == 0:
            cy, cx = rng.choice([5.0, H - 6.0]), rng.choice([4.0, W - 5.0])
        else:
            cy, cx = rng.uniform(0.2 * H, 0.8 * H), rng.uniform(0.2 * W, 0.8 * W)
        a, b = rng.uniform(0.02 * H, 0.3 * H), rng.uniform(0.02 * W, 0.3 * W)
        th = rng.uniform(0, np.pi)
        u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
        v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
        obj = (u / b) ** 2 + (v / a) ** 2 < 1
        obj[int(np.clip(cy, 0, H - 1)), int(np.clip(cx, 0, W - 1))] = True   # never empty
        hy, hx = cy + rng.uniform(-a, a), cx + rng.uniform(-b, b)
        hand = (yy - hy) ** 2 + (xx - hx) ** 2 < rng.uniform(0.04 * H, 0.17 * H) ** 2
        images.append(rng.integers(0, 256, (H, W, 3)).astype(np.uint8))
        objs.append(obj.astype(np.float64) * 255)
        hands.append(hand.astype(np.float64) * 255)
    return images, objs, hands
