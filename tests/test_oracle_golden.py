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
