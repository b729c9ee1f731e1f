"""Host-side candidate gating of the view selection (dynhor_b200.dino_match.select_view) against the statement-by-
statement restatement of pose_initializtion.py:298-321 in oracle/select_view_oracle.py: one designed case per path of
the reference code (argmax, 5- / 10-candidate shortlists, the two 85-degree rejections, the 15 / 30 degree and
max - std fallbacks) plus random cases.  CPU tensors: the gating is plain torch on [N] vectors."""
import math

import numpy as np
import pytest
import torch

from dynhor_b200.dino_match import select_view
from oracle import select_view_oracle as orc


def rz(deg):
    a = torch.as_tensor(deg, dtype=torch.float64) * math.pi / 180.0
    c, s, z, o = torch.cos(a), torch.sin(a), torch.zeros_like(a), torch.ones_like(a)
    return torch.stack([torch.stack([c, -s, z], -1), torch.stack([s, c, z], -1), torch.stack([z, z, o], -1)], -2).float()


def scene(angles, scores, prev_angle):
    """Templates = rotations about z by `angles`; the previous optimised rotation is chosen so that template n is
    |angles[n] - prev_angle| degrees away from it (R_rel = rotations_init @ render_rotations[n])."""
    R = rz(torch.tensor(angles, dtype=torch.float64))
    prev = rz(torch.tensor([-float(prev_angle)], dtype=torch.float64))
    return torch.tensor(scores, dtype=torch.float32), R, prev


def both(cos, R, prev, former, use_former=True):
    k = min(10, len(cos))
    top = torch.topk(cos, k, largest=True).indices
    return select_view(cos, top, R, prev, former_max_idx=former, use_former=use_former), \
        orc.select_view(cos, R, prev, former, use_former)


BASE_ANGLES = [0, 10, 20, 40, 60, 90, 100, 120, 150, 170, 5, 200]
S0 = [.1, .2, .9, .3, .4, .5, .6, .7, .8, .15, .25, .35]
CASES = [
    # name, angles, scores (higher = better), prev angle, former idx, use_former, expected index, expected branch
    ("first_frame", BASE_ANGLES, S0, None, None, True, 2, "argmax"),
    ("use_former_off", BASE_ANGLES, S0, 0, 0, False, 2, "argmax"),
    # 5 best: idx 2 (20 deg), 8 (150), 7 (120), 6 (100), 5 (90); previous pose at 93 -> idx 5 is nearest (3 deg)
    ("top5", BASE_ANGLES, S0, 93, 6, True, 5, "top5"),
    # no former pick: the 10 best now include idx 3 (40 deg); previous pose at 42 -> idx 3
    ("top10", BASE_ANGLES, S0, 42, -1, True, 3, "top10"),
    # previous pose at 270: the ten best are all >= 90 deg away, the nearest view overall (idx 11, 70 deg) is not within 15
    ("far_prev_none_near", BASE_ANGLES, [.1, .2, .9, .3, .4, .5, .6, .7, .8, .15, .25, .05], 270, -1, True, -1,
     "top10+far_prev+none_near"),
    # shortlist {2,8,7,6,5}, previous pose at 3 deg: idx 2 (17 deg) wins the shortlist but is 150 deg from the former pick
    # (idx 9, 170 deg); the nearest view overall (idx 0, 3 deg) is 170 deg from the former pick as well
    ("far_former_then_former_rejects", BASE_ANGLES, S0, 3, 9, True, -1, "top5+far_former+near_rejected_former"),
    # shortlist {5,6,7,8,9} (90..170 deg) all > 85 deg from the previous pose at 2 deg; idx 0 is 2 deg away, 10 deg from the
    # former pick (idx 1) and scores within one standard deviation of the best
    ("far_prev_near_accepted", BASE_ANGLES, [.70, .2, .1, .3, .4, .9, .85, .8, .75, .72, .6, .35], 2, 1, True, 0,
     "top5+far_prev+near_accepted"),
    # the same, but the near view scores far below max - std
    ("far_prev_near_rejected_cos", BASE_ANGLES, [.01, .2, .1, .3, .4, .9, .85, .8, .75, .72, .02, .35], 2, 1, True, -1,
     "top5+far_prev+near_rejected_cos"),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_designed_paths(case):
    name, angles, scores, prev_angle, former, use_former, want, want_branch = case
    if prev_angle is None:
        cos, R, _ = scene(angles, scores, 0.0)
        prev = None
    else:
        cos, R, prev = scene(angles, scores, prev_angle)
    got, (ref, branch) = both(cos, R, prev, former, use_former)
    assert got == ref, (name, got, ref, branch)
    assert ref == want, (name, ref, branch)
    if want_branch is not None:
        assert branch == want_branch, (name, branch)


def test_every_reference_path_is_hit_and_random_cases_agree():
    rng = np.random.default_rng(0)
    seen = set()
    for case in CASES:
        name, angles, scores, prev_angle, former, use_former, _, _ = case
        cos, R, prev = scene(angles, scores, 0.0 if prev_angle is None else prev_angle)
        seen.add(orc.select_view(cos, R, None if prev_angle is None else prev, former, use_former)[1])
    for trial in range(400):
        N = int(rng.integers(10, 40))
        # random rotations: a mix of nearly-aligned clusters (so the 15 / 30 / 85 degree tests all trigger) and wide ones
        spread = rng.choice([5.0, 30.0, 180.0])
        w = rng.normal(size=(N, 3)) * math.radians(spread) / 2
        th = np.linalg.norm(w, axis=1, keepdims=True)
        k = w / np.maximum(th, 1e-12)
        Kx = np.zeros((N, 3, 3))
        Kx[:, 0, 1], Kx[:, 0, 2], Kx[:, 1, 0] = -k[:, 2], k[:, 1], k[:, 2]
        Kx[:, 1, 2], Kx[:, 2, 0], Kx[:, 2, 1] = -k[:, 0], -k[:, 1], k[:, 0]
        Rn = np.eye(3) + np.sin(th)[..., None] * Kx + (1 - np.cos(th))[..., None] * (Kx @ Kx)
        R = torch.from_numpy(Rn).float()
        prev = R[int(rng.integers(N))].T.unsqueeze(0).contiguous() if rng.random() < 0.7 else rz(torch.tensor([rng.uniform(0, 360)]))
        cos = torch.from_numpy(rng.random(N)).float()
        former = int(rng.integers(-1, N))
        got, (ref, branch) = both(cos, R, prev, former)
        assert got == ref, (trial, got, ref, branch)
        seen.add(branch)
    need = {"argmax", "top5", "top10", "top5+far_prev+none_near", "top10+far_prev+none_near",
            "top5+far_former+near_rejected_former", "top5+far_prev+near_accepted", "top5+far_prev+near_rejected_cos"}
    missing = {n for n in need if not any(s == n or s.startswith(n) for s in seen)}
    assert not missing, (missing, sorted(seen))


def test_merge_topk_of_sharded_template_banks():
    """Templates sharded over ranks: merging the per-rank top-k lists equals the top-k of the whole bank, ties going
    to the lowest global index (the kernel's order)."""
    from dynhor_b200.dino_match import merge_topk
    g = torch.Generator().manual_seed(0)
    Fm, k = 7, 5
    sizes = [11, 3, 9, 20]                      # ragged slices, one smaller than k
    scores = torch.rand(Fm, sum(sizes), generator=g)
    scores[:, 12] = scores[:, 30]               # exact ties across ranks
    scores[0, :4] = 2.0                         # and inside one
    vals, idxs, off = [], [], []
    a = 0
    for n in sizes:
        sl = scores[:, a:a + n]
        kk = min(k, n)
        order = torch.argsort(sl, dim=1, descending=True, stable=True)[:, :kk]
        v, i = torch.gather(sl, 1, order), order
        if kk < k:
            v = torch.cat([v, torch.full((Fm, k - kk), float("-inf"))], 1)
            i = torch.cat([i, torch.zeros(Fm, k - kk, dtype=torch.int64)], 1)
        vals.append(v), idxs.append(i), off.append(a)
        a += n
    mv, mi = merge_topk(torch.stack(vals), torch.stack(idxs), off, k)
    order = torch.argsort(scores, dim=1, descending=True, stable=True)[:, :k]
    assert torch.equal(mi, order) and torch.equal(mv, torch.gather(scores, 1, order))
