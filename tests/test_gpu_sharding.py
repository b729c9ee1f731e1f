"""GPU tests of the frame-sharding pieces that run on ONE GPU (the driver's test box has one): shards emulated in one
process (the halo poses and the shared scale's partial gradients moved by hand) must reproduce the single-shard run
bit for bit, with ranges of unequal length; the cost probe; the tick-based mailbox protocol through a loop-back
mailbox.  The 2-GPU versions over NCCL / CUDA-IPC are in test_gpu_multi.py."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import load_golden
from test_gpu_jointopt import _gpu_render_fn, _lw, _model_from_golden, _model_from_seq

pytestmark = pytest.mark.gpu


def _sub(seq, sh):
    B = len(seq["R_init"])
    return {k: (v[sh.start:sh.stop] if isinstance(v, np.ndarray) and len(v) == B and k not in ("verts", "faces")
                else v) for k, v in seq.items()}


def _pose(model, i):
    return torch.cat([model.rotations_object.detach()[i].reshape(6), model.translations_object.detach()[i].reshape(3)])


def _run_emulated(models, fused, iters):
    """One process plays all ranks: before every iteration the boundary poses are copied into the neighbours' halo
    slots; after it the exact partial scale gradients of all shards are handed to every shard (DH_SCALE_DEFERRED)."""
    for _ in range(iters):
        for r, f in enumerate(fused):
            if r > 0:
                f.halo[0].copy_(_pose(models[r - 1], -1))
            if r + 1 < len(fused):
                f.halo[1].copy_(_pose(models[r + 1], 0))
        for f in fused:
            f.run(1, use_graph=True)
        if fused[0].p.optimize_scale:
            parts = torch.stack([f.scale_part for f in fused])
            for f in fused:
                f.apply_scale(parts)


@pytest.fixture(scope="module")
def seq24():
    from dynhor_b200 import synth
    return synth.make_sequence(24, mesh="ico3", seed=9, render_fn=_gpu_render_fn, size=128, period=40)


@pytest.mark.parametrize("bounds", [[0, 12, 24], [0, 5, 6, 24], [0, 1, 9, 17, 24]])
@pytest.mark.parametrize("scale_opt,with_corr", [(False, False), (True, False), (True, True)])
def test_emulated_shards_equal_single(seq24, bounds, scale_opt, with_corr):
    """Any contiguous partition (cost-weighted ranges are ragged) gives the single-shard bits -- poses AND the shared
    object scale (jointopt.py:42-46), whose gradient is summed exactly across the shards; with the correspondence
    term as well (its per-frame sums are formed segment by segment, independently of how many frames share a launch:
    9000 records = 9 tiles = 2 segments per frame here)."""
    from dynhor_b200 import synth
    from dynhor_b200.jointopt import FusedJointOpt
    from dynhor_b200.sharding import FrameShard
    lw = {"lw_sil_obj": 1.0, "lw_smooth_obj": 10.0}
    if with_corr:
        lw["lw_corr_obj"] = 0.01
        seq24 = dict(seq24, correspondences=synth.make_correspondences(seq24, 9000, seed=2, size=128))
    iters, B, lr = 8, 24, 1e-3
    m1 = _model_from_seq(seq24, scale_opt=scale_opt)
    f1 = FusedJointOpt(m1, lw, lr, iters)
    f1.run(iters)
    h1 = f1.history()
    world = len(bounds) - 1
    models, fused = [], []
    for r in range(world):
        sh = FrameShard(r, world, B, bounds=bounds)
        models.append(_model_from_seq(_sub(seq24, sh), scale_opt=scale_opt))
        fused.append(FusedJointOpt(models[-1], lw, lr, iters, shard=sh, keep_sum=f1.keep_sum, exchange=False))
        if with_corr:
            fused[-1].p.corr.w_sum = f1.corr_w_sum      # the sequence-wide sum of weights (an all-reduce in a real run)
    _run_emulated(models, fused, iters)
    assert torch.equal(torch.cat([m.rotations_object.detach() for m in models]), m1.rotations_object.detach())
    assert torch.equal(torch.cat([m.translations_object.detach() for m in models]), m1.translations_object.detach())
    if scale_opt:
        s1 = m1.int_scales_object.detach()
        assert float(s1) != 1.0
        for m in models:
            assert torch.equal(m.int_scales_object.detach(), s1)
    tot = sum(np.asarray(f.history()["loss"]) for f in fused)
    assert np.allclose(tot, h1["loss"], rtol=1e-12)


def test_golden_scale_case_sharded():
    """tests/golden/jointopt_s64_b4_scale (a run of the reference's own jointopt.py with optimize_object_scale=True):
    the sharded path reproduces it like the single-shard path does."""
    from dynhor_b200.jointopt import FusedJointOpt
    from dynhor_b200.sharding import FrameShard
    g = load_golden("s64_b4_scale")
    lw, lr, iters = _lw(g), float(g["lr"]), int(g["iters"])
    B = len(g["rot6d_init"])
    m1 = _model_from_golden(g)
    f1 = FusedJointOpt(m1, lw, lr, iters)
    f1.run(iters)
    models, fused = [], []
    for r in range(2):
        sh = FrameShard(r, 2, B, bounds=[0, 1, B])
        gs = {k: g[k] for k in g.files}
        for k in ("rot6d_init", "trans_init", "K_roi", "target_masks"):
            gs[k] = g[k][sh.start:sh.stop]
        models.append(_model_from_golden(gs))
        fused.append(FusedJointOpt(models[-1], lw, lr, iters, shard=sh, keep_sum=f1.keep_sum, exchange=False))
    _run_emulated(models, fused, iters)
    rot = torch.cat([m.rotations_object.detach() for m in models])
    assert torch.equal(rot, m1.rotations_object.detach())
    assert torch.equal(models[0].int_scales_object.detach(), m1.int_scales_object.detach())
    tot = sum(np.asarray(f.history()["loss"]) for f in fused)
    assert np.allclose(tot, g["ref_loss"], rtol=5e-3)
    assert abs(float(models[1].int_scales_object.detach()) - float(g["ref_final_scale"][0])) < 0.25 * iters * lr


def test_probe_leaves_state_untouched_and_orders_costs(seq24):
    from dynhor_b200.jointopt import FusedJointOpt
    from dynhor_b200.sharding import balanced_bounds, frame_costs_from_blocks
    lw = {"lw_sil_obj": 1.0, "lw_smooth_obj": 10.0}
    m = _model_from_seq(seq24)
    f = FusedJointOpt(m, lw, 1e-4, 4)
    rot0, tr0 = m.rotations_object.detach().clone(), m.translations_object.detach().clone()
    ms = f.probe(4)
    assert ms.shape == (5,) and (ms[:4] > 0).all() and ms[4] == 0.0
    assert torch.equal(m.rotations_object.detach(), rot0) and torch.equal(m.translations_object.detach(), tr0)
    assert int(f.step.item()) == 0
    cost = frame_costs_from_blocks(ms[:4], 0, 24)
    b = balanced_bounds(cost, 3)
    assert b[0] == 0 and b[-1] == 24 and all(y > x for x, y in zip(b, b[1:]))
    # the run after a probe is the run without one
    f.run(3)
    m2 = _model_from_seq(seq24)
    f2 = FusedJointOpt(m2, lw, 1e-4, 4)
    f2.run(3)
    assert torch.equal(m.rotations_object.detach(), m2.rotations_object.detach())


def test_evaluate_after_a_full_run(seq24):
    """evaluate() after max_iters iterations used to return {} (the history had no row left)."""
    from dynhor_b200.jointopt import FusedJointOpt
    lw = {"lw_sil_obj": 1.0, "lw_smooth_obj": 10.0}
    m = _model_from_seq(seq24)
    f = FusedJointOpt(m, lw, 1e-4, 3)
    e0 = f.evaluate()
    f.run(3)
    h = f.history()
    e3 = f.evaluate()
    assert len(h["loss"]) == 3 and h["loss"][0] == e0["loss"][0]
    assert set(e3) == {"loss_smooth_obj", "loss_sil_obj", "iou_object", "loss"} and e3["loss"][0] < e0["loss"][0]
    assert f.history()["loss"] == h["loss"]          # evaluate() does not disturb the history rows


def test_loopback_mailbox_ticks_and_timeout(seq24):
    """The peer-to-peer protocol on one GPU: a 2-rank plan whose "neighbour mailbox" is its own.  Rank 0 of 2 with
    peer_next = own mailbox publishes its last pose into side 0 of itself; we seed side 1 by hand every iteration
    (what rank 1 would publish).  A missing publish trips the wall-clock timeout: the run still terminates and the
    status word is raised instead of a trap."""
    from dynhor_b200 import _lib
    from dynhor_b200.jointopt import FusedJointOpt
    from dynhor_b200.sharding import FrameShard
    lib = _lib.load()
    lw = {"lw_sil_obj": 1.0, "lw_smooth_obj": 10.0}
    sh = FrameShard(0, 2, 24, bounds=[0, 12, 24])
    m = _model_from_seq(_sub(seq24, sh))
    f = FusedJointOpt(m, lw, 1e-4, 8, shard=sh, keep_sum=1000.0, exchange=False, halo_timeout_ms=200)
    mb = ctypes.c_void_p()
    _lib.check(lib.dh_dev_alloc(ctypes.byref(mb), 4 * _lib.MAILBOX_WORDS), "dh_dev_alloc")
    try:
        base = 1000
        f.p.mailbox, f.p.peer_next, f.p.tick_base = mb.value, mb.value, base
        pose = torch.zeros(9, device="cuda")
        pose[0] = pose[3] = 1.0
        pose[8] = 1.8

        def seed(tick):
            slot = mb.value + 4 * ((1 * 4 + (tick & 3)) * 16)
            flag = mb.value + 4 * (128 + 1 * 4 + (tick & 3))
            t = torch.tensor([tick], dtype=torch.int32, device="cuda")
            lib.dh_memcpy_d2d(ctypes.c_void_p(slot), _lib.ptr(pose), 36, _lib.stream_ptr())
            lib.dh_memcpy_d2d(ctypes.c_void_p(flag), _lib.ptr(t), 4, _lib.stream_ptr())
            torch.cuda.synchronize()

        for it in range(3):
            seed(base + it)
            f.run(1, use_graph=False)
        torch.cuda.synchronize()
        f.check_status()
        # what the rank published for its neighbour: its last frame's pose, flagged with the next tick
        words = torch.empty(_lib.MAILBOX_WORDS, dtype=torch.float32, device="cuda")
        lib.dh_memcpy_d2d(_lib.ptr(words), mb, 4 * _lib.MAILBOX_WORDS, _lib.stream_ptr())
        torch.cuda.synchronize()
        tick = base + 3
        got = words[(0 * 4 + (tick & 3)) * 16:(0 * 4 + (tick & 3)) * 16 + 9]
        assert torch.equal(got, _pose(m, -1))
        assert int(words.view(torch.int32)[128 + 0 * 4 + (tick & 3)]) == tick
        # nobody publishes tick base+3 on side 1: the wait gives up after 200 ms, the status word says so
        f.run(1, use_graph=False)
        torch.cuda.synchronize()
        with pytest.raises(_lib.DynhorError, match="timed out"):
            f.check_status()
    finally:
        f.p.mailbox = f.p.peer_next = None
        lib.dh_dev_free(mb)
