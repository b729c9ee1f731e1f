"""GPU: the data path of run.py on either side of the joint optimisation, end to end on the device --
full-frame SAM-style masks -> process_input (run.py:26-72) -> K_roi (pose_initializtion.py:271-277,327) ->
joint_optimize (run.py:155-164).  The full-frame masks are silhouettes of the object under the ground-truth poses,
so the optimisation must pull the perturbed poses back towards them."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_masks_to_target_masks_to_joint_optimisation():
    from dynhor_b200 import synth
    from dynhor_b200.camera import get_K_crop_resize
    from dynhor_b200.jointopt import joint_optimize
    from dynhor_b200.preprocess import process_input_batched
    from dynhor_b200.renderer import Renderer
    B, H, W, S = 6, 512, 512, 256
    verts, faces = synth.icosphere_mesh(3, seed=2)
    R_gt, T_gt = synth.gt_trajectory(B, period=40)
    R0, T0 = synth.perturb_poses(R_gt, T_gt, seed=3)
    K_full = synth.full_frame_K(H, W)
    # full-frame object masks: the object under the ground-truth pose, rendered with the frame's own intrinsics
    K_unit = K_full.copy()
    K_unit[:2] /= H
    vc = torch.from_numpy((verts.astype(np.float64)[None] @ R_gt + T_gt[:, None, :]).astype(np.float32)).cuda()
    frame = Renderer(image_size=H, K=torch.from_numpy(K_unit)[None].cuda(), R=torch.eye(3)[None].cuda(),
                     t=torch.zeros(1, 3).cuda(), orig_size=1, anti_aliasing=False)
    with torch.no_grad():
        sil = frame(vc, torch.from_numpy(faces).cuda()[None].repeat(B, 1, 1), mode="silhouettes")
    obj_masks = (sil > 0.5).to(torch.uint8) * 255                      # SAM convention (run.py:30)
    hand_masks = torch.zeros_like(obj_masks)
    hand_masks[:, 40:130, 180:330] = 255                                # a "hand" across the top edge of the object
    pre = process_input_batched(None, obj_masks, hand_masks)
    target = pre["target_crop_mask"]
    assert set(torch.unique(target).tolist()) <= {-1.0, 0.0, 1.0} and (target == 1).any() and (target == -1).any()
    # pose_initializtion.py:271-277, 327: crop intrinsics from the square box, rows 0-1 over REND_SIZE
    sq = pre["square_bbox"].cpu()
    boxes = torch.stack([sq[:, 0], sq[:, 1], sq[:, 0] + sq[:, 2], sq[:, 1] + sq[:, 2]], 1)
    K_roi = get_K_crop_resize(torch.from_numpy(K_full)[None].repeat(B, 1, 1), boxes, [S])
    K_roi[:, :2] = K_roi[:, :2] / S
    params = [{"rotations": torch.from_numpy(R0[b:b + 1].astype(np.float32)),
               "translations": torch.from_numpy(T0[b:b + 1].astype(np.float32)).reshape(1, 1, 3),
               "K_roi": K_roi[b:b + 1].unsqueeze(0), "target_masks": target[b:b + 1]} for b in range(B)]
    lw = {"lw_sil_obj": 1.0, "lw_smooth_obj": 10.0}
    model, evo = joint_optimize(params, objvertices=verts, objfaces=np.stack([faces] * B), loss_weights=lw,
                                num_iterations=150, lr=1e-3)
    iou = np.asarray(evo["iou_object"])
    assert iou[0] > 0.6 and iou[-1] > iou[0] + 0.03 and iou[-1] > 0.93, (iou[0], iou[-1])
    assert evo["loss_sil_obj"][-1] < 0.5 * evo["loss_sil_obj"][0]


def test_registered_ops_on_the_device():
    """torch.library ops (dynhor_b200/ops.py): opcheck (schema, fake-tensor consistency, autograd registration) with
    real CUDA inputs, and the op path == the module path."""
    from torch.library import opcheck
    import dynhor_b200.ops  # noqa: F401
    from dynhor_b200 import synth
    from dynhor_b200.dino_match import build_bank
    from dynhor_b200.renderer import Renderer
    seq = synth.make_sequence(3, mesh="ico2", seed=1, render_fn=None, size=64)
    vc = torch.from_numpy((seq["verts"].astype(np.float64)[None] @ seq["R_gt"].astype(np.float64)
                           + seq["T_gt"].astype(np.float64)).astype(np.float32)).cuda()
    faces = torch.from_numpy(seq["faces"]).to(torch.int32).cuda()
    K = torch.from_numpy(seq["K_roi"]).cuda()
    args = (vc.requires_grad_(True), faces, K, 64, True, 0.1, 100.0, 1e-4, 1.0)
    opcheck(torch.ops.dynhor.sil_forward.default, args,
            test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    rend = torch.ops.dynhor.sil_forward(*args)
    mod = Renderer(image_size=64, K=K, R=torch.eye(3)[None].cuda(), t=torch.zeros(1, 3).cuda(), orig_size=1)
    assert torch.equal(mod(vc.detach(), faces, mode="silhouettes"), rend.detach()) and float(rend.detach().sum()) > 0
    g = torch.rand_like(rend)
    (gv,) = torch.autograd.grad(rend, vc, g)
    # (the composable backward scatters per-vertex gradients with float atomics: equal up to summation order)
    gv2 = torch.ops.dynhor.sil_backward(vc.detach(), faces, K, g, 64, True, 0.1, 100.0, 1e-4, 1.0)
    assert torch.allclose(gv, gv2, rtol=1e-4, atol=1e-4 * float(gv.abs().max()))
    assert float(gv.abs().sum()) > 0
    d = synth.make_dino_features(40, 6, 24, 64, seed=2, device="cuda")
    fb, tb = build_bank(d["frames"], d["masks"]), build_bank(d["templ"])
    opcheck(torch.ops.dynhor.dino_topk.default, (fb, tb, 5), test_utils=("test_schema", "test_faketensor"))


def test_overlay_of_saved_poses(tmp_path):
    """vis.py:41-55 on the CUDA renderer: the overlay of a saved pose covers the object's mask in the frame."""
    import types
    from dynhor_b200 import poses_io, synth
    from dynhor_b200.renderer import Renderer
    B, H, W = 2, 480, 640
    verts, faces = synth.icosphere_mesh(3, seed=2)
    R, T = synth.gt_trajectory(B, period=40)
    m = types.SimpleNamespace(rotations_object=torch.from_numpy(np.ascontiguousarray(R[:, :, :2])).float().cuda(),
                              translations_object=torch.from_numpy(T).float().reshape(B, 1, 3).cuda())
    paths = ["rgb/%04d.jpg" % i for i in range(B)]
    poses_io.save_obj_infos(m, synth.full_frame_K(H, W), paths, str(tmp_path))
    infos = poses_io.load_obj_infos(str(tmp_path), paths)
    frames = np.full((B, H, W, 3), 200, np.uint8)
    out = poses_io.overlay_mesh(frames, infos, verts, faces)
    changed = (out != frames).any(-1)
    # ground truth: where the projected vertices fall
    K = synth.full_frame_K(H, W).astype(np.float64)
    for b in range(B):
        uv = (verts.astype(np.float64) @ R[b] + T[b]) @ K.T
        uv = uv[:, :2] / uv[:, 2:3]
        inside = changed[b][np.clip(uv[:, 1].round().astype(int), 0, H - 1), np.clip(uv[:, 0].round().astype(int), 0, W - 1)]
        assert inside.mean() > 0.9 and 0.01 < changed[b].mean() < 0.5
