"""The pose files run.py writes and vis.py reads (dynhor_b200/poses_io.py): layout, names and values (CPU)."""
import os
import types

import numpy as np
import torch

from dynhor_b200 import poses_io, synth


def _model(B, seed=0):
    R, T = synth.gt_trajectory(B, period=40)
    m = types.SimpleNamespace()
    m.rotations_object = torch.nn.Parameter(torch.from_numpy(np.ascontiguousarray(R[:, :, :2])).float())   # geometry.py:38
    m.translations_object = torch.nn.Parameter(torch.from_numpy(T).float().reshape(B, 1, 3))
    return m, R, T


def test_obj_infos_round_trip(tmp_path):
    B = 5
    model, R, T = _model(B)
    paths = ["/data/seq/rgb/%06d.jpg" % (3 * i) for i in range(B)]
    K = synth.full_frame_K(480, 640)
    written = poses_io.save_obj_infos(model, K, paths, str(tmp_path))
    assert [os.path.basename(p) for p in written] == ["%06d.npz" % (3 * i) for i in range(B)]
    d = np.load(written[2])
    assert sorted(d.files) == ["K", "R", "T"] and d["R"].shape == (3, 3) and d["T"].shape == (1, 3) and d["K"].shape == (3, 3)
    # R on disk is object -> camera: the TRANSPOSE of the matrix whose first two columns are the 6D parameters (run.py:166)
    assert np.allclose(d["R"], R[2].T, atol=1e-6) and np.allclose(d["T"], T[2].reshape(1, 3), atol=1e-7)
    assert np.array_equal(d["K"], K)
    infos = poses_io.load_obj_infos(str(tmp_path), paths + ["/data/seq/rgb/999999.jpg"])
    assert infos[-1] is None and all(i is not None for i in infos[:-1]) and infos[0]["obj_scale"] == 1.0
    # what vis.py:52 computes from a file equals the optimiser's own transform v @ R6d + T
    v = np.random.default_rng(0).normal(size=(7, 3)).astype(np.float32)
    assert np.allclose(v @ infos[1]["R"].T + infos[1]["T"], v @ R[1] + T[1], atol=1e-5)
