"""CPU: the builder-defined correspondence term -- kernel arithmetic (host build of dh_core.h::corr_record) and the
launch plan of k_corr against the oracle (oracle/corr_oracle.py).  No GPU needed."""
import ctypes

import numpy as np
import pytest
import torch

import emu_lib
from dynhor_b200 import synth
from helpers import rel_err
from oracle import corr_oracle
from oracle.jointopt_oracle import rot6d_to_matrix


def _case(B=3, C=777, seed=5, size=64):
    seq = synth.make_sequence(B, 120, 160, mesh="ico2", seed=seed, render_fn=None, size=size)
    rec = synth.make_correspondences(seq, C, seed=seed, size=size, outliers=0.1)
    return seq, rec


def _oracle_sums_and_grads(seq, rec, size, delta, scale=1.0):
    r6 = torch.tensor(seq["rot6d_init"], requires_grad=True)
    T = torch.tensor(seq["T_init"], requires_grad=True)
    s = torch.tensor([scale], requires_grad=True)
    sums = corr_oracle.corr_frame_sums(torch.from_numpy(rec), rot6d_to_matrix(r6), T, s.abs(),
                                       torch.from_numpy(seq["K_roi"]), size, delta)
    R = rot6d_to_matrix(r6).detach()
    Rl = R.clone().requires_grad_(True)
    sums2 = corr_oracle.corr_frame_sums(torch.from_numpy(rec), Rl, T, s.abs(), torch.from_numpy(seq["K_roi"]),
                                        size, delta)
    gR, gT, gs = torch.autograd.grad(sums2.sum(), [Rl, T, s])
    return sums.detach().numpy(), R.numpy(), gR.numpy(), gT.numpy(), gs.numpy()


@pytest.mark.parametrize("delta", [0.5, 2.0])
@pytest.mark.parametrize("scale", [1.0, -0.8])
def test_corr_record_arithmetic_vs_oracle(delta, scale):
    size = 64
    seq, rec = _case(size=size)
    sums_o, R, gR_o, gT_o, gs_o = _oracle_sums_and_grads(seq, rec, size, delta, scale)
    sums = emu_lib.corr_frames(rec, R, seq["T_init"], abs(scale), seq["K_roi"], size, delta).astype(np.float64)
    assert rel_err(sums[:, 12], sums_o) < 1e-5
    assert rel_err(sums[:, 0:3], gT_o.reshape(-1, 3)) < 1e-4
    assert rel_err(abs(scale) * sums[:, 3:12], gR_o.reshape(-1, 9)) < 1e-4
    gs = np.sign(scale) * (R.reshape(-1, 9) * sums[:, 3:12]).sum()
    assert abs(gs - gs_o[0]) <= 1e-4 * abs(gs_o[0]) + 1e-6
    # both Huber branches are exercised
    assert 0 < (sums_o > 0).sum()


def test_zero_weight_and_exact_hit_records_contribute_nothing():
    size = 64
    seq, rec = _case(B=2, C=10, size=size)
    rec[:, :, 5] = 0.0
    s = emu_lib.corr_frames(rec, seq["R_init"], seq["T_init"], 1.0, seq["K_roi"], size, 1.0)
    assert np.all(s == 0.0)


SEG = 8   # kSegTiles (dh_corr.cu): tiles per segment, the unit one CTA always sums as a whole


def _cut(c, T, G, tpf):
    """dh_corr.cu::corr_cut: CTA c's first tile = the equal cut c*T/G moved forward to the next segment start."""
    t = c * T // G
    b, k = divmod(t, tpf)
    ks = -(-k // SEG) * SEG
    return b * tpf + min(ks, tpf)


@pytest.mark.parametrize("B,C,sms", [(300, 10000, 148), (4, 100, 148), (512, 50000, 148), (1, 5000, 148), (7, 2050, 2),
                                      (64, 10000, 148), (3, 2, 148), (2048, 50000, 148), (1907, 50000, 148)])
def test_plan_covers_every_tile_once_and_never_splits_a_segment(B, C, sms):
    """Every record tile is streamed by exactly one CTA, and a segment (8 consecutive tiles of one frame, slot =
    segment number) is never cut: a frame's partial sums are formed the same way whatever B and the grid are, which
    is what makes a frame-sharded run reproduce the single-GPU bits."""
    from dynhor_b200.corr import plan
    p = plan(B, C, sms)
    G, nslots, tpf = p["grid"], p["nslots"], p["tiles_per_frame"]
    assert tpf == -(-C // 1024) and nslots == -(-tpf // SEG) and 1 <= G <= max(1, 3 * sms)
    T = B * tpf
    seen = np.zeros(T, int)
    owner = {}
    for i in range(G):
        t0, t1 = _cut(i, T, G, tpf), _cut(i + 1, T, G, tpf)
        assert t1 >= t0
        seen[t0:t1] += 1
        for t in range(t0, t1):
            b, k = divmod(t, tpf)
            assert owner.setdefault((b, k // SEG), i) == i      # one CTA per segment
    assert np.all(seen == 1) and len(owner) == B * nslots
    if G > 1:   # balanced up to one segment
        sizes = [_cut(i + 1, T, G, tpf) - _cut(i, T, G, tpf) for i in range(G)]
        assert max(sizes) - min(sizes) <= 2 * SEG


def test_synthetic_correspondences_are_consistent_with_the_ground_truth():
    seq, rec = _case(B=4, C=500, size=64)
    s = corr_oracle.corr_frame_sums(torch.from_numpy(rec), torch.from_numpy(seq["R_gt"]),
                                    torch.from_numpy(seq["T_gt"]), torch.ones(1), torch.from_numpy(seq["K_roi"]),
                                    64, 1.0)
    s0 = corr_oracle.corr_frame_sums(torch.from_numpy(rec), torch.from_numpy(seq["R_init"]),
                                     torch.from_numpy(seq["T_init"]), torch.ones(1), torch.from_numpy(seq["K_roi"]),
                                     64, 1.0)
    assert float(s.sum()) < float(s0.sum())     # the ground-truth pose explains the matches better than the init
    assert rec.shape == (4, 500, 6) and rec.dtype == np.float32 and np.all(rec[..., 5] > 0)
