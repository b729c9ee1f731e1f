"""Where the end-to-end time of joint_optimize goes (host buffers in, poses out): synchronised phase times.
    python tools/e2e_breakdown.py [frames] [iters] [corr]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["DH_TIMING"] = "1"
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from dynhor_b200.jointopt import joint_optimize  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
C = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
seq = bench.make_range(0, B, B, C)
params, _ = bench.host_parameters(seq, 0, B, C)
faces_b = np.stack([seq["faces"]] * B)
lw = bench.loss_weights(C)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model, evo = joint_optimize(params, objvertices=seq["verts"], objfaces=faces_b, loss_weights=lw, num_iterations=iters,
                                lr=1e-4)
    rot = model.rotations_object.detach().cpu()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    print(f"rep {rep}: total {dt:.1f} ms; " + ", ".join(f"{k} {v:.1f}" for k, v in model.timing))
