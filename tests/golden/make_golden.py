"""Generate tests/golden/*.npz by running the REFERENCE'S OWN PYTHON on CPU in the build container.

Run from the repo root (needs /root/reference, which does not exist on the GPU box -- the committed .npz files
are what travels):

    python tests/golden/make_golden.py

What runs unmodified from /root/reference/ObjTracker: jointopt.py (Joint_Optimizer, joint_optimize),
utils/losses.py (Losses, batch_mask_iou), utils/geometry.py, utils/camera.py.
What is substituted, and why:
  * `neural_renderer` (requirements.txt:8, third-party CUDA-only, not installed, not vendored) -> the CPU
    restatement oracle/nr_oracle.py + oracle/nmr_oracle.c.  The rasteriser internals therefore stay
    "parity unpinned"; everything around them is the reference's code.
  * `.cuda()` calls (jointopt.py:56,104-105; losses.py:38-39,67) -> no-ops, `torch.cuda.FloatTensor`
    (camera.py:26, evaluated at import) -> torch.FloatTensor: the container has no GPU driver.
  * utils.losses.REND_SIZE is patched for the small cases (constants.py:2 is 256).
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/ObjTracker"
sys.path.insert(0, ROOT)

from oracle import nr_oracle, jointopt_oracle  # noqa: E402
from dynhor_b200 import synth  # noqa: E402


def import_reference():
    torch.cuda.FloatTensor = torch.FloatTensor
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    nr = nr_oracle.as_neural_renderer_module()
    sys.modules["neural_renderer"] = nr
    sys.modules["neural_renderer.renderer"] = nr.renderer
    sys.path.insert(0, REF)
    import jointopt as ref_jointopt  # noqa
    import utils.losses as ref_losses  # noqa
    import utils.geometry as ref_geometry  # noqa
    import utils.camera as ref_camera  # noqa
    return ref_jointopt, ref_losses, ref_geometry, ref_camera


class _Board:
    def add_scalar(self, *a, **k):
        pass


def oracle_render_fn(vc, faces, K, size):
    B = len(vc)
    r = nr_oracle.Renderer(image_size=size, K=torch.from_numpy(K), R=torch.eye(3)[None], t=torch.zeros(1, 3),
                           orig_size=1, anti_aliasing=False)
    return r(torch.from_numpy(vc), torch.from_numpy(faces)[None].repeat(B, 1, 1), mode="silhouettes").numpy()


def run_case(name, ref, B, mesh, size, iters, lr, lw, seed, occluder=True, scale_opt=False):
    ref_jointopt, ref_losses, _, _ = ref
    ref_losses.REND_SIZE = size
    seq = synth.make_sequence(B, mesh=mesh, seed=seed, render_fn=oracle_render_fn, size=size, occluder=occluder)
    params = synth.to_object_parameters(seq)
    faces_b = np.stack([seq["faces"]] * B)

    # (1) one forward/backward of the reference model by hand -> gradients at iteration 0
    model = ref_jointopt.Joint_Optimizer(
        translations_object=torch.cat([p["translations"] for p in params]),
        rotations_object=torch.cat([p["rotations"] for p in params]),
        verts_object_og=torch.from_numpy(seq["verts"]),
        faces_object=torch.from_numpy(faces_b),
        camintr_rois_object=torch.cat([p["K_roi"][:, 0] for p in params]),
        target_masks_object=torch.cat([p["target_masks"] for p in params]),
        int_scale_init=1, optimize_object_scale=scale_opt)
    loss_dict, metric_dict = model(loss_weights=lw)
    loss = sum(loss_dict[k] * lw[k.replace("loss", "lw")] for k in loss_dict)
    loss.backward()
    g_rot = model.rotations_object.grad.detach().numpy().copy()
    g_trans = model.translations_object.grad.detach().numpy().copy()
    g_scale = model.int_scales_object.grad.detach().numpy().copy() if scale_opt else np.zeros(1, np.float32)
    with torch.no_grad():
        verts0 = model.get_verts_object()
        rend0 = model.losses.sil_renderer(verts0, model.faces_object, mode="silhouettes").numpy()

    # (2) oracle-side rasteriser maps at iteration 0 (oracle-defined; regression pin for the oracle itself)
    with torch.no_grad():
        faces2 = torch.cat((model.faces_object, model.faces_object.flip(-1)), dim=1)
        proj = nr_oracle.projection(verts0, model.camintr_rois_object, torch.eye(3)[None], torch.zeros(1, 3),
                                    torch.zeros(1, 5), 1)
        fv = nr_oracle.vertices_to_faces(proj, faces2).numpy()
    maps = nr_oracle.rasterize_forward_np(fv, size * 2)

    # (3) the reference loop, unmodified
    _, evo = ref_jointopt.joint_optimize(
        object_parameters=params, objvertices=seq["verts"], objfaces=faces_b, loss_weights=lw,
        num_iterations=iters, lr=lr, board=_Board(), optimize_object_scale=scale_opt)
    model2, _ = None, None
    # rerun to fetch final parameters (joint_optimize returns the model)
    model2, evo2 = ref_jointopt.joint_optimize(
        object_parameters=params, objvertices=seq["verts"], objfaces=faces_b, loss_weights=lw,
        num_iterations=iters, lr=lr, board=_Board(), optimize_object_scale=scale_opt)
    assert evo["loss"] == evo2["loss"], "reference run is not deterministic on CPU"

    # (4) the restated oracle must reproduce the reference's numbers (same torch ops, same order)
    orc = jointopt_oracle.JointOptOracle(seq["rot6d_init"], seq["T_init"], seq["verts"], seq["faces"], seq["K_roi"],
                                         seq["target_masks"], lr=lr, image_size=size,
                                         optimize_object_scale=scale_opt)
    evo_o = orc.run(lw, iters)
    for k in ("loss", "loss_sil_obj", "loss_smooth_obj", "iou_object"):
        a, b = np.asarray(evo[k]), np.asarray(evo_o[k])
        assert np.allclose(a, b, rtol=1e-5, atol=1e-8), (name, k, a, b)
    assert np.allclose(orc.rotations_object.detach().numpy(), model2.rotations_object.detach().numpy(), atol=1e-6)

    out = dict(
        verts=seq["verts"], faces=seq["faces"].astype(np.int32), K_roi=seq["K_roi"],
        target_masks=seq["target_masks"].astype(np.int8), rot6d_init=seq["rot6d_init"], trans_init=seq["T_init"],
        size=np.int32(size), iters=np.int32(iters), lr=np.float64(lr),
        lw_sil_obj=np.float64(lw["lw_sil_obj"]), lw_smooth_obj=np.float64(lw["lw_smooth_obj"]),
        scale_opt=np.int32(scale_opt),
        ref_grad_rot6d=g_rot, ref_grad_trans=g_trans, ref_grad_scale=g_scale,
        ref_rend0=rend0.astype(np.float32),
        ref_loss=np.asarray(evo["loss"], np.float64),
        ref_loss_sil=np.asarray(evo["loss_sil_obj"], np.float64),
        ref_loss_smooth=np.asarray(evo["loss_smooth_obj"], np.float64),
        ref_iou=np.asarray(evo["iou_object"], np.float64),
        ref_final_rot6d=model2.rotations_object.detach().numpy(),
        ref_final_trans=model2.translations_object.detach().numpy(),
        ref_final_scale=model2.int_scales_object.detach().numpy(),
        orc_face_index0=maps["face_index"], orc_alpha0=np.packbits(maps["alpha"] > 0.5, axis=-1),
    )
    path = os.path.join(ROOT, "tests", "golden", f"jointopt_{name}.npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB; loss", evo["loss"][0], "->", evo["loss"][-1],
          "iou", evo["iou_object"][0], "->", evo["iou_object"][-1])


def geometry_case(ref):
    """rot6d_to_matrix / projection / transform of the reference on seeded inputs."""
    _, _, ref_geometry, ref_camera = ref
    g = torch.Generator().manual_seed(0)
    rot6d = torch.randn(7, 3, 2, generator=g)
    R = ref_geometry.rot6d_to_matrix(rot6d)
    assert torch.equal(R, jointopt_oracle.rot6d_to_matrix(rot6d))
    verts = torch.randn(7, 64, 3, generator=g) * 0.3 + torch.tensor([0.0, 0.0, 1.7])
    K = torch.eye(3)[None].repeat(7, 1, 1)
    K[:, 0, 0] = 1.9 + torch.rand(7, generator=g)
    K[:, 1, 1] = 2.1 + torch.rand(7, generator=g)
    K[:, 0, 2] = 0.5 + 0.1 * torch.randn(7, generator=g)
    K[:, 1, 2] = 0.5 + 0.1 * torch.randn(7, generator=g)
    proj = ref_camera.projection(verts, K, torch.eye(3)[None], torch.zeros(1, 3), 1, torch.zeros(1, 5))
    assert torch.equal(proj, nr_oracle.projection(verts, K, torch.eye(3)[None], torch.zeros(1, 3),
                                                  torch.zeros(1, 5), 1))
    T = torch.randn(7, 1, 3, generator=g)
    vt = ref_camera.compute_transformation_persp(verts[0], T, R, torch.ones(1) * 1.3)
    assert torch.equal(vt, jointopt_oracle.transform_verts(verts[0], T, R, torch.ones(1) * 1.3))
    path = os.path.join(ROOT, "tests", "golden", "geometry.npz")
    np.savez_compressed(path, rot6d=rot6d.numpy(), R=R.numpy(), verts=verts.numpy(), K=K.numpy(),
                        proj=proj.numpy(), T=T.numpy(), verts_t=vt.numpy())
    print("geometry ->", path)


def import_pose_initializtion():
    for name, attrs in {
        "pytorch3d": [], "pytorch3d.renderer": ["PerspectiveCameras", "RasterizationSettings", "MeshRenderer",
                                                "MeshRasterizer", "SoftPhongShader", "SoftSilhouetteShader",
                                                "BlendParams"],
        "pytorch3d.structures": ["Meshes"], "detectron2": [], "detectron2.structures": ["BitMasks", "BoxMode"],
        "detectron2.layers": ["ROIAlign"], "detectron2.structures.boxes": ["BoxMode"],
        "detectron2.layers.roi_align": ["ROIAlign"],
    }.items():
        m = types.ModuleType(name)
        for a in attrs:
            setattr(m, a, type(a, (), {"__init__": lambda self, *a, **k: None}))
        sys.modules.setdefault(name, m)
    import pose_initializtion as ref_pi  # noqa
    return ref_pi


def stage1_case():
    """ObjTracker.coarse_forward + the mode="coarse" loop of find_optimal_pose (pose_initializtion.py:143-155,
    346-358), reference code unmodified.  pose_initializtion.py imports pytorch3d / detectron2 names at module
    level (:17-30); they are only used by the textured `forward` path, so empty stand-in modules are enough."""
    ref_pi = import_pose_initializtion()
    B, size = 1, 128
    seq = synth.make_sequence(3, mesh="ico3", seed=6, render_fn=oracle_render_fn, size=size)
    f = 1  # one frame, like the reference's per-frame loop
    K = torch.from_numpy(seq["K_roi"][f:f + 1].copy())
    rot6d = torch.from_numpy(seq["rot6d_init"][f:f + 1].copy())
    trans = torch.from_numpy(seq["T_init"][f:f + 1].copy())
    model = ref_pi.ObjTracker(ref_image=seq["target_masks"][f], vertices=torch.from_numpy(seq["verts"]),
                              faces=torch.from_numpy(seq["faces"])[None], textures=None, dino_model=None,
                              gt_dino_feat=torch.zeros(1), rotation_init=rot6d, translation_init=trans,
                              num_initializations=1, K=K)
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    losses, ious = [], []
    g0 = None
    for it in range(6):
        opt.zero_grad()
        loss_dict, iou = model.coarse_forward()
        total = sum(loss_dict.values()).sum()
        total.backward()
        if it == 0:
            g0 = (model.rotations.grad.numpy().copy(), model.translations.grad.numpy().copy())
        opt.step()
        losses.append(float(total.detach()))
        ious.append(iou.numpy().copy())
    path = os.path.join(ROOT, "tests", "golden", "stage1_coarse.npz")
    np.savez_compressed(path, verts=seq["verts"], faces=seq["faces"].astype(np.int32), K_roi=K.numpy(),
                        target_mask=seq["target_masks"][f].astype(np.int8), rot6d_init=rot6d.numpy(),
                        trans_init=trans.numpy(), lr=np.float64(0.01), ref_loss=np.asarray(losses),
                        ref_iou=np.asarray(ious), ref_grad_rot=g0[0], ref_grad_trans=g0[1],
                        ref_final_rot=model.rotations.detach().numpy(),
                        ref_final_trans=model.translations.detach().numpy())
    print("stage1_coarse ->", path, "loss", losses[0], "->", losses[-1], "iou", ious[0], "->", ious[-1])


def stage1_multi_case():
    """The same reference code with num_initializations = 4 (one ObjTracker, four candidate poses of one frame,
    pose_initializtion.py:329-360), an occluder in the target mask, and one candidate pushed partly out of the view so
    that the off-screen penalty (:119-141, weight 100000 at :154) and its gradient are exercised."""
    ref_pi = import_pose_initializtion()
    size, n = 128, 4     # (not 3: geometry.py:24 calls torch.cross without dim, which picks the batch axis at B = 3)
    seq = synth.make_sequence(n, mesh="ico3", seed=8, render_fn=oracle_render_fn, size=size, occluder=True)
    f = 0
    K = torch.from_numpy(seq["K_roi"][f:f + 1].copy())
    # four candidates around frame f's pose: its perturbed initial pose and the neighbouring frames' rotations
    rot6d = torch.from_numpy(seq["rot6d_init"].copy())
    trans = torch.from_numpy(seq["T_init"][f:f + 1].copy()).repeat(n, 1, 1)
    trans[3, 0, 0] += 0.45      # slides a part of the object out of the view: off-screen penalty > 0
    model = ref_pi.ObjTracker(ref_image=seq["target_masks"][f], vertices=torch.from_numpy(seq["verts"]),
                              faces=torch.from_numpy(seq["faces"])[None], textures=None, dino_model=None,
                              gt_dino_feat=torch.zeros(1), rotation_init=rot6d, translation_init=trans,
                              num_initializations=n, K=K)
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    losses, ious, per_init = [], [], []
    g0 = None
    for it in range(5):
        opt.zero_grad()
        loss_dict, iou = model.coarse_forward()
        lv = sum(loss_dict.values())
        total = lv.sum()
        total.backward()
        if it == 0:
            g0 = (model.rotations.grad.numpy().copy(), model.translations.grad.numpy().copy())
            off0 = loss_dict["offscreen"].detach().numpy().copy()
        opt.step()
        losses.append(float(total.detach()))
        per_init.append(lv.detach().numpy().copy())
        ious.append(iou.numpy().copy())
    print('offscreen0', off0)
    assert off0[3] > 0 and off0[0] == 0
    path = os.path.join(ROOT, "tests", "golden", "stage1_multi.npz")
    np.savez_compressed(path, verts=seq["verts"], faces=seq["faces"].astype(np.int32), K_roi=K.numpy(),
                        target_mask=seq["target_masks"][f].astype(np.int8), rot6d_init=rot6d.numpy(),
                        trans_init=trans.numpy(), lr=np.float64(0.01), ref_loss=np.asarray(losses),
                        ref_loss_per_init=np.asarray(per_init), ref_offscreen0=off0,
                        ref_iou=np.asarray(ious), ref_grad_rot=g0[0], ref_grad_trans=g0[1],
                        ref_final_rot=model.rotations.detach().numpy(),
                        ref_final_trans=model.translations.detach().numpy())
    print("stage1_multi ->", path, "loss", losses[0], "->", losses[-1], "offscreen0", off0, "iou", ious[0], "->", ious[-1])


def main():
    if "--stage1-multi" in sys.argv:   # added in round 2; the other fixtures are not regenerated
        import_reference()
        stage1_multi_case()
        return
    ref = import_reference()
    stage1_case()
    geometry_case(ref)
    lw = {"lw_sil_obj": 1.0, "lw_smooth_obj": 10.0}  # configs/custom_shoes.yaml:17-18
    run_case("s64_b5", ref, B=5, mesh="ico3", size=64, iters=8, lr=1e-4, lw=lw, seed=1)
    run_case("s256_b4", ref, B=4, mesh="ico2", size=256, iters=3, lr=1e-4, lw=lw, seed=0)
    run_case("s128_b6_lr", ref, B=6, mesh="ico3", size=128, iters=6, lr=1e-3,
             lw={"lw_sil_obj": 2.0, "lw_smooth_obj": 3.0}, seed=2, occluder=True)
    run_case("s64_b4_scale", ref, B=4, mesh="ico2", size=64, iters=5, lr=1e-3, lw=lw, seed=3, scale_opt=True)


if __name__ == "__main__":
    main()
