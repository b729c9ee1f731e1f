import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import test_gpu_jointopt as T
from helpers import rel_err
from dynhor_b200 import synth
from dynhor_b200.jointopt import FusedJointOpt
from oracle import jointopt_oracle as jo
seq = synth.make_sequence(4, mesh="uv50x100", seed=7, render_fn=T._oracle_render_fn, period=300)
lw = {"lw_sil_obj": 1.0, "lw_smooth_obj": 10.0}
model = T._model_from_seq(seq)
fused = FusedJointOpt(model, lw, 1e-4, 4)
g_rot, g_tr, _ = fused.grads()
orc = jo.JointOptOracle(seq["rot6d_init"], seq["T_init"], seq["verts"], seq["faces"], seq["K_roi"], seq["target_masks"], lr=1e-4)
out, grads = orc.loss_and_grads(lw)
for b in range(4):
    print(b, "rot", rel_err(g_rot[b].cpu().numpy(), grads["rot6d"][b]), "trans", rel_err(g_tr[b].cpu().numpy(), grads["trans"][b]))
