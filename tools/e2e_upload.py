"""Host-side pieces of joint_optimize's model build, timed one by one (synchronised): the per-key uploads, the module,
the fused plan.    python tools/e2e_upload.py [frames] [corr]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from dynhor_b200 import jointopt as J  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
C = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
seq = bench.make_range(0, B, B, C)
params, _ = bench.host_parameters(seq, 0, B, C)
lw = bench.loss_weights(C)
verts = J.tensorify(seq["verts"]).cuda()
faces = J._SharedFaces(np.stack([seq["faces"]] * B))


def timed(label, fn, reps=4):
    out = None
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{label:28s} " + " ".join(f"{t:7.2f}" for t in ts) + " ms")
    return out


trans = timed("translations", lambda: J._stack_frames(params, "translations"))
rots = timed("rotations", lambda: J._stack_frames(params, "rotations"))
K = timed("K_roi", lambda: J._stack_frames(params, "K_roi", pick=lambda t: t[:, 0]))
masks = timed("target_masks", lambda: J._stack_frames(params, "target_masks", dtype=torch.float32))
corr = timed("correspondences", lambda: J._stack_frames(params, "correspondences")) if C else None


def foreach(key):
    ts = [p[key] for p in params]
    out = torch.empty((len(ts),) + tuple(ts[0].shape[1:]), dtype=ts[0].dtype, device="cuda")
    torch._foreach_copy_(list(out.split(1)), ts, non_blocking=True)
    return out


timed("masks via _foreach_copy_", lambda: foreach("target_masks"))
if C:
    timed("corr via _foreach_copy_", lambda: foreach("correspondences"))
model = timed("Joint_Optimizer", lambda: J.Joint_Optimizer(
    translations_object=trans, rotations_object=rots, verts_object_og=verts, faces_object=faces.dev,
    target_masks_object=masks, camintr_rois_object=K, int_scale_init=1, optimize_object_scale=False,
    correspondences=corr, corr_delta=1.0))


def fused():
    f = J.FusedJointOpt(model, lw, 1e-4, 20, corr_on=C > 0)
    f.release()
    return f


timed("FusedJointOpt", fused)
