"""CPU: the arithmetic the CUDA kernels run (dynhor_b200/csrc/dh_core.h, built for the host by tests/emu) against
the golden vectors and the oracle.  Coverage / face ownership bit-exact, losses 1e-4, gradients 1e-3 -- the same
bars the GPU tests apply to the real kernels."""
import numpy as np
import pytest
import torch

import emu_lib as E
from helpers import JOINT_CASES, golden_alpha, load_golden, rel_err, unpack_alpha
from oracle import nr_oracle


@pytest.mark.parametrize("name", JOINT_CASES)
def test_emu_full_iteration_vs_golden(name):
    g = load_golden(name)
    S = int(g["size"])
    out = E.full_grads(g["verts"], g["faces"], g["K_roi"], g["target_masks"], g["rot6d_init"], g["trans_init"], S,
                       float(g["lw_sil_obj"]), float(g["lw_smooth_obj"]))
    assert np.array_equal(out["fidx"], g["orc_face_index0"])            # face ownership: bit-exact
    assert np.array_equal(unpack_alpha(out["abits"]), golden_alpha(g))  # coverage: bit-exact
    assert np.array_equal(out["rend"], g["ref_rend0"])                  # pooled silhouette: bit-exact
    assert abs(out["loss_sil_obj"] - g["ref_loss_sil"][0]) <= 1e-4 * g["ref_loss_sil"][0]
    assert abs(out["loss_smooth_obj"] - g["ref_loss_smooth"][0]) <= 1e-4 * g["ref_loss_smooth"][0]
    assert abs(out["iou_object"] - g["ref_iou"][0]) <= 1e-6
    assert rel_err(out["grad_rot6d"], g["ref_grad_rot6d"]) < 1e-3
    assert rel_err(out["grad_trans"], g["ref_grad_trans"]) < 1e-3
    if int(g["scale_opt"]):
        assert abs(out["grad_scale"] - float(g["ref_grad_scale"][0])) < 1e-3 * abs(float(g["ref_grad_scale"][0]))


def _random_scene(seed, B=3, S=64, n_rings=8, n_seg=12):
    from dynhor_b200 import synth
    verts, faces = synth.uv_sphere_mesh(n_rings, n_seg, seed)
    seq = synth.make_sequence(B, mesh=(verts, faces), seed=seed, size=S)
    return seq


@pytest.mark.parametrize("seed,aa", [(0, True), (1, True), (2, False)])
def test_emu_backward_arbitrary_grad_vs_oracle(seed, aa):
    """API-mode backward: random upstream gradient on the rendered image; per-face-vertex gradients must agree
    with the oracle's edge-scan backward run on the SAME projected faces and maps."""
    S = 64
    seq = _random_scene(seed, S=S)
    B = len(seq["R_init"])
    cam = (seq["verts"][None].astype(np.float64) @ seq["R_init"].astype(np.float64) + seq["T_init"]).astype(np.float32)
    proj = E.project_cam(cam, seq["K_roi"])
    is_ = 2 * S if aa else S
    fidx, abits = E.raster(proj, seq["faces"], is_)
    faces2 = np.concatenate([seq["faces"], seq["faces"][:, ::-1]], 0)
    fv = proj[:, :, :3][np.arange(B)[:, None, None], faces2[None]]          # [B,2F,3,3]
    maps = nr_oracle.rasterize_forward_np(fv, is_)
    assert np.array_equal(maps["face_index"], fidx)
    rng = np.random.default_rng(seed)
    g_rend = rng.normal(size=(B, S, S)).astype(np.float32)
    g_rend[rng.random(size=g_rend.shape) < 0.3] = 0.0
    # oracle: avg-pool backward + un-flip, then the edge scan
    g_t = torch.from_numpy(g_rend)
    if aa:
        g512 = torch.repeat_interleave(torch.repeat_interleave(g_t, 2, 1), 2, 2) * 0.25
    else:
        g512 = g_t
    g512 = g512.flip(1).contiguous().numpy()
    gf_o = nr_oracle.rasterize_backward_np(fv, maps["face_index"], maps["alpha"], g512)[..., :2]
    pos, neg = E.grad_signs(g_rend)
    gf_e, gv = E.backward(proj, seq["faces"], fidx, abits, g_rend, pos.reshape(B, S, -1), neg.reshape(B, S, -1), S,
                          aa, verts_cam=cam, K=seq["K_roi"])
    assert np.allclose(gf_e, gf_o, rtol=1e-5, atol=1e-9)
    assert np.abs(gf_o).max() > 0
    # vertex gradients vs torch autograd through the oracle projection
    vt = torch.from_numpy(cam).requires_grad_(True)
    p = nr_oracle.projection(vt, torch.from_numpy(seq["K_roi"]), torch.eye(3)[None], torch.zeros(1, 3),
                             torch.zeros(1, 5), 1)
    f = nr_oracle.vertices_to_faces(p, torch.from_numpy(faces2)[None].repeat(B, 1, 1))
    gfo3 = np.zeros(fv.shape, np.float32)
    gfo3[..., :2] = gf_o
    f.backward(torch.from_numpy(gfo3))
    assert rel_err(gv, vt.grad.numpy()) < 1e-4


def test_emu_projection_bit_exact_vs_torch():
    g = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "geometry.npz"))
    assert np.array_equal(E.rot6d_to_R(g["rot6d"]), g["R"])
    proj, cam = E.project_pose(g["verts"][0], g["R"], g["T"], 1.3, g["K"])
    assert np.array_equal(cam, g["verts_t"])
    p2 = E.project_cam(g["verts"], g["K"])
    assert np.array_equal(p2[:, :, :3], g["proj"])


def test_emu_adam_vs_torch():
    rng = np.random.default_rng(0)
    p0 = rng.normal(size=37).astype(np.float32)
    pt = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([pt], lr=1e-3)
    p, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    for t in range(1, 8):
        gr = (rng.normal(size=37) * (10.0 ** rng.integers(-3, 2))).astype(np.float32)
        pt.grad = torch.from_numpy(gr.copy())
        opt.step()
        E.adam(p, gr, m, v, 1e-3, t)
        assert np.allclose(p, pt.detach().numpy(), rtol=0, atol=2e-7), t


def test_emu_smoothness_closed_form_sharded_equals_unsharded():
    """Frame sharding: gradients of a 2-shard split with halos equal the single-shard gradients."""
    seq = _random_scene(3, B=7)
    mom = E.mesh_moments(seq["verts"])
    V = len(seq["verts"])
    st = E.smooth_terms(seq["rot6d_init"], seq["T_init"], 1.0, mom, V, 7, 10.0)
    pose = np.concatenate([seq["rot6d_init"].reshape(7, 6), seq["T_init"].reshape(7, 3)], 1).astype(np.float32)
    a = E.smooth_terms(seq["rot6d_init"][:4], seq["T_init"][:4], 1.0, mom, V, 7, 10.0, None, pose[4])
    b = E.smooth_terms(seq["rot6d_init"][4:], seq["T_init"][4:], 1.0, mom, V, 7, 10.0, pose[3], None)
    assert np.array_equal(np.concatenate([a, b]), st)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_emu_raster_triangle_soup_vs_bruteforce_oracle(seed):
    """Edge cases of the binned / span-filtered rasteriser against the brute-force oracle: slivers, degenerate
    (collinear, repeated-vertex) faces, faces partly or fully off-screen, behind the near plane, exact depth ties,
    vertices exactly on pixel centres."""
    rng = np.random.default_rng(seed)
    V, F, is_ = 60, 150, 64
    proj = np.zeros((1, V, 4), np.float32)
    proj[0, :, 0] = rng.uniform(-1.3, 1.3, V)
    proj[0, :, 1] = rng.uniform(-1.3, 1.3, V)
    proj[0, :, 2] = rng.uniform(0.05, 3.0, V)
    # a few vertices exactly on pixel centres / with identical depth (ties)
    proj[0, :8, 0] = (2 * rng.integers(0, is_, 8) + 1 - is_) / is_
    proj[0, :8, 1] = (2 * rng.integers(0, is_, 8) + 1 - is_) / is_
    proj[0, 8:20, 2] = 1.0
    faces = rng.integers(0, V, size=(F, 3)).astype(np.int32)
    faces[:10, 2] = faces[:10, 1]                      # repeated vertex
    for k in range(10, 25):                            # slivers: third vertex almost on the edge
        a, b = proj[0, faces[k, 0]], proj[0, faces[k, 1]]
        t = rng.uniform(-0.2, 1.2)
        proj[0, V - 1 - (k - 10), :2] = a[:2] + t * (b[:2] - a[:2]) + rng.normal(size=2) * 1e-6
        faces[k, 2] = V - 1 - (k - 10)
    fidx, abits = E.raster(proj, faces, is_)
    faces2 = np.concatenate([faces, faces[:, ::-1]], 0)
    fv = proj[:, :, :3][np.arange(1)[:, None, None], faces2[None]]
    maps = nr_oracle.rasterize_forward_np(fv, is_)
    assert np.array_equal(maps["face_index"], fidx)
    assert np.array_equal(unpack_alpha(abits), maps["alpha"] > 0.5)
    assert (fidx >= 0).mean() > 0.2


def test_emu_real_shoe_mesh_vs_oracle():
    """The reference's own object prior (tests/golden/shoe_mesh.npz <- assets/shoes, normalised like run.py:110-112):
    a non-convex real mesh through the kernel arithmetic -- face ownership / coverage bit-exact against the oracle's
    brute-force rasteriser, losses and pose gradients against the oracle pipeline."""
    import os
    from helpers import GOLDEN
    from dynhor_b200 import synth
    from oracle import jointopt_oracle as jo
    m = np.load(os.path.join(GOLDEN, "shoe_mesh.npz"))
    verts, faces = m["verts"].astype(np.float32), m["faces"].astype(np.int64)
    assert verts.shape == (2502, 3) and faces.shape == (5000, 3)
    S = 64

    def render_fn(vc, fc, K, size):
        B = len(vc)
        r = nr_oracle.Renderer(image_size=size, K=torch.from_numpy(K), R=torch.eye(3)[None], t=torch.zeros(1, 3),
                               orig_size=1, anti_aliasing=False)
        return r(torch.from_numpy(vc), torch.from_numpy(fc)[None].repeat(B, 1, 1), mode="silhouettes").numpy()

    seq = synth.make_sequence(2, mesh=(verts, faces), seed=5, size=S, render_fn=render_fn, period=5)
    mt = np.where(seq["target_masks"] > 0, 1, np.where(seq["target_masks"] >= 0, 0, -1)).astype(np.int8)
    out = E.full_grads(verts, faces.astype(np.int32), seq["K_roi"], mt, seq["rot6d_init"], seq["T_init"], S, 1.0, 10.0)
    B = 2
    faces2 = np.concatenate([faces, faces[:, ::-1]], 0)
    maps = nr_oracle.rasterize_forward_np(out["proj"][:, :, :3][np.arange(B)[:, None, None], faces2[None]], 2 * S)
    assert np.array_equal(out["fidx"], maps["face_index"])
    assert np.array_equal(unpack_alpha(out["abits"]), maps["alpha"] > 0.5)
    orc = jo.JointOptOracle(seq["rot6d_init"], seq["T_init"], verts, faces, seq["K_roi"], seq["target_masks"], lr=1e-4,
                            image_size=S)
    ref, grads = orc.loss_and_grads({"lw_sil_obj": 1.0, "lw_smooth_obj": 10.0})
    assert abs(out["loss_sil_obj"] - ref["loss_sil_obj"]) <= 1e-4 * ref["loss_sil_obj"]
    assert abs(out["iou_object"] - ref["iou_object"]) <= 1e-6
    assert rel_err(out["grad_rot6d"], grads["rot6d"]) < 1e-3
    assert rel_err(out["grad_trans"], grads["trans"]) < 1e-3


@pytest.mark.parametrize("case", ["mesh_sized", "mixed", "few_lanes", "giant", "degenerate"])
def test_emu_span_list_hands_out_every_crossing_once(case):
    """The backward's one-crossing-per-lane enumeration (k_backward<lists>: compacted span list + mark / popcount step,
    helpers in dh_core.h shared with the kernel) visits exactly the crossings of the plain nested loops, in the same
    order -- also when a batch has more than 2^16 crossings (32 faces across the whole 512-pixel raster: the span
    starts are kept modulo 2^16) and when lanes or spans are empty."""
    import ctypes
    rng = np.random.default_rng({"mesh_sized": 1, "mixed": 2, "few_lanes": 3, "giant": 4, "degenerate": 5}[case])
    is_ = 512
    for trial in range(40 if case != "giant" else 4):
        n_have = 32
        c = rng.uniform(-20, is_ + 20, (32, 1, 2))
        if case == "mesh_sized":
            tri = c + rng.uniform(-4, 4, (32, 3, 2))
        elif case == "mixed":
            tri = c + rng.uniform(-1, 1, (32, 3, 2)) * rng.choice([0.3, 5.0, 60.0, 700.0], (32, 1, 1))
        elif case == "few_lanes":
            n_have = int(rng.integers(0, 6))
            tri = c + rng.uniform(-30, 30, (32, 3, 2))
        elif case == "giant":
            # a triangle's spans add up to at most ~2 x (width + height) of the raster = 2049 scan lines: 32 such faces
            # give the largest batch there can be, 65 568 crossings -- just past 2^16
            big = np.array([[806.152244, 575.274903], [448.0696, -0.0979789455], [-87.0836405, -200.319568]])
            tri = np.stack([big if trial == 0 else big + rng.uniform(-0.01, 0.01, (3, 2)) for _ in range(32)])
        else:                      # axis-aligned and zero-length edges, vertices on pixel centres
            tri = np.round(c + rng.uniform(-6, 6, (32, 3, 2)))
            tri[::3, 1] = tri[::3, 0]
        px = np.ascontiguousarray(tri[:, :, 0], np.float32)
        py = np.ascontiguousarray(tri[:, :, 1], np.float32)
        cap = 32 * 6 * is_
        flat = np.full((cap, 3), -7, np.int32)
        ref = np.full((cap, 3), -9, np.int32)
        fn = E.lib().emu_span_list
        fn.restype = ctypes.c_int
        T = fn(E._p(px), E._p(py), n_have, is_, E._p(flat), E._p(ref), cap)
        assert T >= 0
        if case == "giant" and trial == 0:
            assert T == 32 * 2049 > 1 << 16
        if case == "few_lanes" and n_have == 0:
            assert T == 0
        assert np.array_equal(flat[:T], ref[:T])
        assert (ref[:T, 2] >= 0).all() and (ref[:T, 2] < is_).all()
