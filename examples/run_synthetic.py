#!/usr/bin/env python
"""Stage 2 of ObjTracker/run.py (run.py:152-179) on a synthetic sequence, with the drop-in joint_optimize.

    python examples/run_synthetic.py --config_path configs/custom_shoes.yaml [--frames 64]
    torchrun --nproc-per-node 8 examples/run_synthetic.py --config_path configs/custom_shoes.yaml --frames 1000

Reads the same YAML keys as run.py:94-97,152-153,162, calls joint_optimize with run.py:155-164's arguments and
writes exps/<seq>/<exp>/obj_infos/<frame>.npz with R, T, K exactly like run.py:166-179 (the on-disk contract vis.py
reads).  Stage 1 (per-frame initialisation) is replaced by perturbed ground-truth poses and masks rendered with the
CUDA silhouette renderer."""
import argparse
import os
import sys

import numpy as np
import torch
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from dynhor_b200 import poses_io, synth  # noqa: E402
from dynhor_b200.geometry import rot6d_to_matrix  # noqa: E402
from dynhor_b200.jointopt import joint_optimize  # noqa: E402
from dynhor_b200.renderer import Renderer  # noqa: E402


def render_fn(vc, faces, K, size):
    B = len(vc)
    r = Renderer(image_size=size, K=torch.from_numpy(K).cuda(), R=torch.eye(3)[None].cuda(),
                 t=torch.zeros(1, 3).cuda(), orig_size=1, anti_aliasing=False)
    with torch.no_grad():
        return r(torch.from_numpy(vc).cuda(), torch.from_numpy(faces).cuda()[None].repeat(B, 1, 1),
                 mode="silhouettes").cpu().numpy()


class Board:
    """Stand-in for tensorboardX.SummaryWriter (run.py:127) that keeps the scalars in memory."""

    def __init__(self):
        self.scalars = []

    def add_scalar(self, k, v, step):
        self.scalars.append((k, v, step))


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--config_path", type=str, required=True)
    parser.add_argument("--frames", type=int, default=64)
    parser.add_argument("--height", type=int, default=480)
    parser.add_argument("--width", type=int, default=640)
    args = parser.parse_args()
    with open(args.config_path, "r") as f:
        config = yaml.safe_load(f)
    if "RANK" in os.environ:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    rank = int(os.environ.get("RANK", "0"))
    seq = synth.make_sequence(args.frames, args.height, args.width, mesh="uv50x100", seed=0, render_fn=render_fn)
    object_parameters = synth.to_object_parameters(seq)
    obj_faces = seq["faces"]
    sample_folder = os.path.join("exps", config["seq_name"], config["exp_name"])
    os.makedirs(sample_folder, exist_ok=True)
    board = Board()
    num_iterations = config["system"]["joint_num_iterations"]
    loss_weights = config["system"]["loss"]
    model, loss_evolution = joint_optimize(
        object_parameters=object_parameters,
        objvertices=seq["verts"],
        objfaces=np.stack([obj_faces for _ in range(args.frames)]),
        optimize_object_scale=False,
        loss_weights=loss_weights,
        num_iterations=num_iterations,
        lr=config["system"]["joint_lr"],
        board=board,
    )
    camintr = synth.full_frame_K(args.height, args.width)
    if rank == 0:
        # run.py:166-179: one obj_infos/<frame id>.npz per frame (R object -> camera, T, K)
        poses_io.save_obj_infos(model, camintr, ["rgb/{:06d}.jpg".format(i) for i in range(args.frames)], sample_folder)
        err0 = np.abs(seq["R_init"] - seq["R_gt"]).max()
        err1 = np.abs(rot6d_to_matrix(model.rotations_object).detach().cpu().numpy() - seq["R_gt"]).max()
        print(f"{args.frames} frames, {num_iterations} iterations: loss {loss_evolution['loss'][0]:.5f} -> "
              f"{loss_evolution['loss'][-1]:.5f}, IoU {loss_evolution['iou_object'][0]:.4f} -> "
              f"{loss_evolution['iou_object'][-1]:.4f}, max |R - R_gt| {err0:.4f} -> {err1:.4f}; "
              f"poses in {sample_folder}/obj_infos/")
    if "RANK" in os.environ:
        torch.distributed.destroy_process_group()
