"""The on-disk contract on the output side of the joint optimisation and the overlay that reads it.

    save_obj_infos   run.py:166-179   one `obj_infos/<frame id>.npz` per frame with
                                      R [3,3] = rot6d_to_matrix(rotations_object)[i]^T  (object -> camera, :166)
                                      T [1,3] = translations_object[i], K [3,3] = the full-frame intrinsics
    load_obj_infos   vis.py:43-52     reads them back (R, T, optional obj_scale), skipping missing frames like vis.py
    overlay_mesh     vis.py:41-55     the optimised mesh drawn over the frames.  The reference shades it with pyrender
                                      (OSMesa); here the CUDA silhouette renderer gives the coverage of
                                      (obj_scale * verts) @ R^T + T under the full-frame camera and the overlay is an
                                      alpha blend of a flat colour -- enough to check poses by eye without pyrender.
"""
import os

import numpy as np
import torch

from .geometry import _rot6d_to_matrix_torch, rot6d_to_matrix


def frame_id(image_path):
    """run.py:177: file name without directory and 4-character extension."""
    return image_path.split("/")[-1][:-4]


def pose_arrays(model):
    """(R [B,3,3] object -> camera, T [B,1,3]) as numpy, from a model returned by joint_optimize (run.py:166-170)."""
    rot = model.rotations_object.detach()
    R = (rot6d_to_matrix(rot) if rot.is_cuda else _rot6d_to_matrix_torch(rot)).transpose(1, 2)
    return R.cpu().numpy(), model.translations_object.detach().cpu().numpy()


def save_obj_infos(model, camintr, image_paths, sample_folder):
    """run.py:166-179.  Returns the written paths."""
    R, T = pose_arrays(model)
    if len(image_paths) != len(R):
        raise ValueError(f"{len(image_paths)} image paths for {len(R)} optimised frames")
    out_dir = os.path.join(sample_folder, "obj_infos")
    os.makedirs(out_dir, exist_ok=True)
    paths = []
    for i, image_path in enumerate(image_paths):
        path = os.path.join(out_dir, "{}.npz".format(frame_id(image_path)))
        np.savez(path, R=R[i], T=T[i], K=np.asarray(camintr))
        paths.append(path)
    return paths


def load_obj_infos(sample_folder, image_paths):
    """-> list of dicts {R, T, K, obj_scale} (None where a frame has no file, which vis.py:44 skips)."""
    out = []
    for image_path in image_paths:
        path = os.path.join(sample_folder, "obj_infos", "{}.npz".format(frame_id(image_path)))
        if not os.path.exists(path):
            out.append(None)
            continue
        d = np.load(path)
        out.append({"R": d["R"], "T": d["T"], "K": d["K"] if "K" in d.files else None,
                    "obj_scale": float(d["obj_scale"]) if "obj_scale" in d.files else 1.0})
    return out


def overlay_mesh(images, infos, verts_can, faces, focal=None, color=(0.2, 0.6, 1.0), alpha=0.6, render_size=512):
    """images: list / array of [H,W,3] uint8 frames; infos: load_obj_infos output.  Returns [n,H,W,3] uint8 with the
    mesh's coverage under the camera of vis.py:38,53 (focal 1.2 * min(H, W), principal point (W//2, H//2)) blended in."""
    from .renderer import Renderer
    imgs = np.stack([np.asarray(im) for im in images])
    n, H, W = imgs.shape[:3]
    side = max(H, W)
    f = float(focal) if focal is not None else 1.2 * min(H, W)
    # unit-image intrinsics of the side x side square whose top-left corner is the image origin
    K = torch.tensor([[f / side, 0.0, (W // 2) / side], [0.0, f / side, (H // 2) / side], [0.0, 0.0, 1.0]])
    keep = [i for i, info in enumerate(infos) if info is not None]
    out = imgs.copy()
    if not keep:
        return out
    v = torch.as_tensor(np.asarray(verts_can), dtype=torch.float32)
    cam = torch.stack([(infos[i]["obj_scale"] * v) @ torch.as_tensor(infos[i]["R"], dtype=torch.float32).T
                       + torch.as_tensor(infos[i]["T"], dtype=torch.float32).reshape(1, 3) for i in keep]).cuda()
    r = Renderer(image_size=render_size, K=K[None].cuda(), R=torch.eye(3)[None].cuda(), t=torch.zeros(1, 3).cuda(),
                 orig_size=1, anti_aliasing=False)
    with torch.no_grad():
        sil = r(cam, torch.as_tensor(np.asarray(faces)).cuda()[None].expand(len(keep), -1, -1), mode="silhouettes")
        cov = torch.nn.functional.interpolate(sil[:, None], size=(side, side), mode="nearest")[:, 0, :H, :W]
    cov = cov.cpu().numpy()[..., None] * alpha
    col = (np.asarray(color, np.float32) * 255.0).reshape(1, 1, 3)
    for j, i in enumerate(keep):
        out[i] = np.clip(imgs[i] * (1.0 - cov[j]) + col * cov[j], 0, 255).astype(np.uint8)
    return out
