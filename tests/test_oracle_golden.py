"""CPU: the oracle (oracle/) reproduces the golden vectors generated from the reference's own Python
(tests/golden/make_golden.py).  Pins the oracle; nothing here touches the product."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN, JOINT_CASES, golden_alpha, load_golden
from oracle import jointopt_oracle as jo
from oracle import nr_oracle

import os


def test_geometry_golden():
    g = np.load(os.path.join(GOLDEN, "geometry.npz"))
    R = jo.rot6d_to_matrix(torch.from_numpy(g["rot6d"]))
    assert np.array_equal(R.numpy(), g["R"])
    proj = nr_oracle.projection(torch.from_numpy(g["verts"]), torch.from_numpy(g["K"]), torch.eye(3)[None],
                                torch.zeros(1, 3), torch.zeros(1, 5), 1)
    assert np.array_equal(proj.numpy(), g["proj"])
    vt = jo.transform_verts(torch.from_numpy(g["verts"][0]), torch.from_numpy(g["T"]), torch.from_numpy(g["R"]),
                            torch.ones(1) * 1.3)
    assert np.array_equal(vt.numpy(), g["verts_t"])


@pytest.mark.parametrize("name", JOINT_CASES)
def test_jointopt_oracle_matches_reference_run(name):
    g = load_golden(name)
    lw = {"lw_sil_obj": float(g["lw_sil_obj"]), "lw_smooth_obj": float(g["lw_smooth_obj"])}
    orc = jo.JointOptOracle(g["rot6d_init"], g["trans_init"], g["verts"], g["faces"].astype(np.int64), g["K_roi"],
                            g["target_masks"].astype(np.float32), lr=float(g["lr"]), image_size=int(g["size"]),
                            optimize_object_scale=bool(g["scale_opt"]))
    with torch.no_grad():
        rend = orc.render().numpy()
    assert np.array_equal(rend, g["ref_rend0"])
    out, grads = orc.loss_and_grads(lw)
    assert np.allclose(grads["rot6d"], g["ref_grad_rot6d"], rtol=1e-5, atol=1e-7)
    assert np.allclose(grads["trans"], g["ref_grad_trans"], rtol=1e-5, atol=1e-7)
    evo = orc.run(lw, int(g["iters"]))
    assert np.allclose(evo["loss"], g["ref_loss"], rtol=1e-6)
    assert np.allclose(evo["loss_sil_obj"], g["ref_loss_sil"], rtol=1e-6)
    assert np.allclose(evo["loss_smooth_obj"], g["ref_loss_smooth"], rtol=1e-6)
    assert np.allclose(evo["iou_object"], g["ref_iou"], rtol=1e-6)
    assert np.allclose(orc.rotations_object.detach().numpy(), g["ref_final_rot6d"], atol=1e-6)
    assert np.allclose(orc.translations_object.detach().numpy(), g["ref_final_trans"], atol=1e-6)


@pytest.mark.parametrize("name", JOINT_CASES[:2])
def test_oracle_rasteriser_maps_pinned(name):
    """Regression pin of the (parity-unpinned) C rasteriser restatement: maps stored at generation time."""
    g = load_golden(name)
    B, size = len(g["rot6d_init"]), int(g["size"])
    R = jo.rot6d_to_matrix(torch.from_numpy(g["rot6d_init"]))
    verts = jo.transform_verts(torch.from_numpy(g["verts"]), torch.from_numpy(g["trans_init"]), R, torch.ones(1))
    faces = torch.from_numpy(g["faces"].astype(np.int64))[None].repeat(B, 1, 1)
    faces2 = torch.cat((faces, faces.flip(-1)), dim=1)
    proj = nr_oracle.projection(verts, torch.from_numpy(g["K_roi"]), torch.eye(3)[None], torch.zeros(1, 3),
                                torch.zeros(1, 5), 1)
    maps = nr_oracle.rasterize_forward_np(nr_oracle.vertices_to_faces(proj, faces2).numpy(), size * 2)
    assert np.array_equal(maps["face_index"], g["orc_face_index0"])
    assert np.array_equal(maps["alpha"] > 0.5, golden_alpha(g))


@pytest.mark.parametrize("name", ["stage1_coarse", "stage1_multi"])
def test_stage1_oracle_vs_reference_run(name):
    """oracle/stage1_oracle.py against the runs of the reference's own pose_initializtion.ObjTracker.coarse_forward +
    Adam loop stored by tests/golden/make_golden.py (one candidate; four candidates with an occluder and one of
    them partly off-screen)."""
    import os
    from helpers import GOLDEN
    from oracle import stage1_oracle
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    n = len(g["rot6d_init"])
    orc = stage1_oracle.Stage1Oracle(np.repeat(g["target_mask"][None].astype(np.float32), n, 0), g["verts"],
                                     g["faces"].astype(np.int64), g["rot6d_init"], g["trans_init"], g["K_roi"],
                                     lr=float(g["lr"]))
    for it in range(len(g["ref_loss"])):
        lv, iou, off, grads = orc.step()
        assert np.isclose(float(lv.sum()), g["ref_loss"][it], rtol=1e-6)
        assert np.allclose(iou.numpy(), g["ref_iou"][it], rtol=1e-6)
        if it == 0:
            assert np.allclose(grads[0].numpy(), g["ref_grad_rot"], rtol=1e-5, atol=1e-7 * np.abs(g["ref_grad_rot"]).max())
            assert np.allclose(grads[1].numpy(), g["ref_grad_trans"], rtol=1e-5,
                               atol=1e-7 * np.abs(g["ref_grad_trans"]).max())
            if "ref_offscreen0" in g.files:
                assert np.allclose(OFF * off.numpy(), g["ref_offscreen0"], rtol=1e-6)
    assert np.allclose(orc.rotations.detach().numpy(), g["ref_final_rot"], atol=1e-6)
    assert np.allclose(orc.translations.detach().numpy(), g["ref_final_trans"], atol=1e-6)


OFF = 100000.0
