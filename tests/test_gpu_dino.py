"""GPU parity tests of the DINO template matcher (tcgen05 GEMM + fused split-K reduction / top-k) against the CPU
oracle (oracle/dino_oracle.py = pose_initializtion.py:295-296,309 verbatim).  Top-k indices bit-exact, scores
within 2e-3 absolute (bf16 banks, fp32 accumulation; scores are O(1) cosines)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SCORE_ATOL = 2e-3


def _run(N, Fm, P, D, k, seed=0):
    from dynhor_b200 import synth
    from dynhor_b200.dino_match import build_bank, dino_cos_topk
    from oracle import dino_oracle
    d = synth.make_dino_features(N, Fm, P, D, seed=seed)
    s_o, v_o, i_o = dino_oracle.dino_cos_topk(d["frames"], d["masks"], d["templ"], k)
    tb = build_bank(d["templ"].cuda())
    fb = build_bank(d["frames"].cuda(), d["masks"].cuda())
    s_g, v_g, i_g = dino_cos_topk(fb, tb, k)
    torch.cuda.synchronize()
    s_g, v_g, i_g = s_g.cpu(), v_g.cpu(), i_g.cpu()
    err = float((s_g - s_o).abs().max())
    assert err <= SCORE_ATOL, err
    # the selection itself: exactly torch.topk of the kernel's own (tie-free) scores
    v_t, i_t = torch.topk(s_g, k, dim=1, largest=True)
    assert torch.equal(i_g, i_t) and torch.equal(v_g, v_t)
    # against the oracle's indices wherever top-k is well-posed: consecutive oracle scores further apart than
    # twice the score tolerance cannot swap
    srt = torch.sort(s_o, dim=1, descending=True).values
    kk = min(k, N - 1)
    gap_ok = (srt[:, :kk] - srt[:, 1:kk + 1]).min(dim=1).values > 2 * max(err, 1e-6) * 1.5
    assert torch.equal(i_g[gap_ok], i_o[gap_ok])
    print(f"N={N} Fm={Fm} K={P * D} k={k}: max |score err| {err:.2e}, well-posed frames {float(gap_ok.float().mean()):.2f}")
    assert torch.allclose(v_g, v_o, atol=SCORE_ATOL, rtol=0)
    # planted best match is rank 0
    assert torch.equal(i_g[:, 0], d["match"])
    # values are the scores at the returned indices, in non-increasing order
    assert torch.equal(v_g, torch.gather(s_g, 1, i_g))
    assert bool((v_g[:, :-1] >= v_g[:, 1:]).all()) if k > 1 else True
    return s_g


@pytest.mark.parametrize("N,Fm,P,D,k", [
    (200, 40, 37, 64, 5),      # one tile each way, K = 2368 (37 k-blocks)
    (130, 7, 35, 72, 10),      # ragged M tile, K = 2520 (partial last k-block)
    (300, 330, 16, 64, 5),     # two frame tile pairs (Fm > 320), ragged
    (1000, 300, 49, 96, 10),   # BASELINE configs[3] tile counts (8 x 1 tiles), reduced K
    (64, 1, 8, 8, 1),          # single frame, argmax only, K = 64 (a single k-block)
])
def test_dino_topk_vs_oracle(N, Fm, P, D, k):
    _run(N, Fm, P, D, k)


def test_dino_full_size_properties():
    """BASELINE configs[3]: 1k templates x 300 frames, ViT-S/14 tokens (P=1369, D=384), k=10.  The CPU oracle
    needs ~0.5 s per frame here, so 6 frames are checked against it and the rest through properties: planted match
    at rank 0, scores of a frame against its own template bank entry near 1/(1+noise^2)^0.5, run-to-run
    determinism, linearity in the frame bank."""
    from dynhor_b200 import synth
    from dynhor_b200.dino_match import build_bank, dino_cos_topk
    from oracle import dino_oracle
    N, Fm, P, D, k = 1000, 300, 1369, 384, 10
    d = synth.make_dino_features(N, Fm, P, D, seed=3, device="cuda")
    tb = build_bank(d["templ"])
    fb = build_bank(d["frames"], d["masks"])
    s1, v1, i1 = dino_cos_topk(fb, tb, k)
    s2, v2, i2 = dino_cos_topk(fb, tb, k)
    assert torch.equal(s1, s2) and torch.equal(i1, i2)
    assert torch.equal(i1[:, 0].cpu(), d["match"])
    sub = [0, 1, 2, 150, 298, 299]
    s_o, v_o, i_o = dino_oracle.dino_cos_topk(d["frames"][sub].cpu(), d["masks"][sub].cpu(), d["templ"].cpu(), k)
    assert torch.allclose(s1[sub].cpu(), s_o, atol=SCORE_ATOL, rtol=0)
    assert torch.equal(i1[sub, :1].cpu(), i_o[:, :1])
    # linearity: scoring 0.5 * frames halves every score (exact in bf16: a power-of-two scaling)
    s3, _, i3 = dino_cos_topk((fb.float() * 0.5).bfloat16(), tb, k)
    assert torch.equal(s3, s1 * 0.5) and torch.equal(i3, i1)


def test_select_view_matches_reference_gating():
    """Host-side candidate gating (pose_initializtion.py:298-321) on top of the kernel's top-k."""
    from dynhor_b200.dino_match import select_view
    from dynhor_b200.synth import axis_angle_to_matrix
    g = torch.Generator().manual_seed(0)
    N = 50
    R = torch.from_numpy(axis_angle_to_matrix(np.random.default_rng(0).normal(size=(N, 3)))).float().cuda()
    cos = torch.rand(N, generator=g).cuda()
    top = torch.topk(cos, 10).indices
    assert select_view(cos, top, R, None) == int(top[0])
    prev = R[int(top[2])].T.unsqueeze(0).contiguous()
    idx = select_view(cos, top, R, prev, former_max_idx=int(top[2]))
    assert idx == int(top[2])


@pytest.mark.parametrize("cluster", ["1", "2", "8"])
def test_dino_cluster_variants(cluster, monkeypatch):
    """The multicast GEMM with other cluster sizes than the default 4 (tuning knob DH_DINO_CLUSTER), including a
    template count that needs padded tiles (N = 1100 -> 9 tiles -> 10/12/16 with clusters of 2/4/8)."""
    monkeypatch.setenv("DH_DINO_CLUSTER", cluster)
    _run(1100, 50, 24, 64, 5, seed=5)


def test_fp32_rescoring_restores_the_reference_order_on_near_ties():
    """bf16 banks can swap candidates whose fp32 scores are closer than ~2e-3.  Templates 1..4 are copies of
    template 0 with a perturbation small enough that their fp32 scores differ by ~1e-4: the kernel's order among them
    is arbitrary at bf16 precision, rescore_topk_fp32 must return the oracle's."""
    from dynhor_b200.dino_match import build_bank, dino_cos_topk, rescore_topk_fp32
    from oracle import dino_oracle
    g = torch.Generator().manual_seed(4)
    N, Fm, P, D, k = 64, 6, 40, 64, 5
    templ = torch.nn.functional.normalize(torch.randn(N, P, D, generator=g), dim=-1)
    frames = torch.nn.functional.normalize(templ[:Fm] + 0.3 * torch.randn(Fm, P, D, generator=g), dim=-1)
    for f in range(Fm):                       # k near-duplicates of every frame's best template
        for j in range(1, k):
            templ[8 + f * k + j] = torch.nn.functional.normalize(templ[f] + 2e-3 * j * torch.randn(P, D, generator=g), dim=-1)
    masks = (torch.rand(Fm, P, generator=g) < 0.7).float()
    masks[:, 0] = 1
    s_o, v_o, i_o = dino_oracle.dino_cos_topk(frames, masks, templ, k)
    gaps = (v_o[:, :-1] - v_o[:, 1:]).min()
    assert 0 < float(gaps) < 2e-3                                  # the near-ties are there, below bf16 resolution
    _, v, i = dino_cos_topk(build_bank(frames.cuda(), masks.cuda()), build_bank(templ.cuda()), k)
    assert torch.equal(torch.sort(i.cpu(), 1).values, torch.sort(i_o, 1).values)   # same candidate SET
    v2, i2 = rescore_topk_fp32(frames.cuda(), masks.cuda(), templ, i)               # templates stay on the host
    assert torch.equal(i2.cpu(), i_o)
    assert torch.allclose(v2.cpu(), v_o, atol=1e-6)
