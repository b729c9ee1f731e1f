"""CPU, world_size 2 over gloo: the host side of frame sharding (dynhor_b200/sharding.py) -- range partition, the
per-iteration one-frame pose halo exchange, the keep-mask all-reduce and the final pose gather -- driving the
emulated per-frame arithmetic, must reproduce the single-shard result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import emu_lib as E
from dynhor_b200.sharding import FrameShard, allgather_frames, allreduce_sum_, exchange_halo


def test_frame_shard_partition():
    for B, W in [(300, 8), (7, 2), (1000, 8), (13, 4), (4096, 8), (8, 8)]:
        shards = [FrameShard(r, W, B) for r in range(W)]
        assert shards[0].start == 0 and shards[-1].stop == B
        assert all(a.stop == b.start for a, b in zip(shards, shards[1:]))
        assert max(s.B for s in shards) - min(s.B for s in shards) <= 1
        assert not shards[0].has_prev and not shards[-1].has_next
    with pytest.raises(ValueError):
        FrameShard(0, 4, 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dynhor_b200 import synth
        verts, faces = synth.uv_sphere_mesh(8, 12, 5)
        seq = synth.make_sequence(B, mesh=(verts, faces), seed=5, size=64)
        sh = FrameShard(rank, world, B)
        rot = torch.from_numpy(seq["rot6d_init"][sh.start:sh.stop].copy())
        tr = torch.from_numpy(seq["T_init"][sh.start:sh.stop].copy())
        mom = E.mesh_moments(verts)
        halo = torch.zeros(2, 9)
        edge = torch.zeros(2, 9)
        keep = torch.tensor([float(100 + rank)], dtype=torch.float64)
        allreduce_sum_(keep, sh)
        hist = []
        for it in range(3):
            edge[0, :6], edge[0, 6:] = rot[0].reshape(6), tr[0].reshape(3)
            edge[1, :6], edge[1, 6:] = rot[-1].reshape(6), tr[-1].reshape(3)
            exchange_halo(edge[0], edge[1], sh, halo[0], halo[1])
            st = E.smooth_terms(rot.numpy(), tr.numpy(), 1.0, mom, len(verts), B, 10.0,
                                halo[0].numpy() if sh.has_prev else None, halo[1].numpy() if sh.has_next else None)
            # a deterministic "optimiser step" driven by the smoothness gradient only
            g6 = E.rot6d_backward(rot.numpy(), st[:, 3:12])
            rot = rot - 0.05 * torch.from_numpy(g6).float()
            tr = tr - 0.05 * torch.from_numpy(st[:, 0:3]).float().reshape(-1, 1, 3)
            part = torch.tensor([st[:, 13].sum()], dtype=torch.float64)
            allreduce_sum_(part, sh)
            hist.append(float(part))
        rot_all = allgather_frames(rot, sh)
        tr_all = allgather_frames(tr, sh)
        if rank == 0:
            q.put((rot_all.numpy(), tr_all.numpy(), hist, float(keep)))
    finally:
        dist.destroy_process_group()


def _single(B):
    from dynhor_b200 import synth
    verts, faces = synth.uv_sphere_mesh(8, 12, 5)
    seq = synth.make_sequence(B, mesh=(verts, faces), seed=5, size=64)
    rot = torch.from_numpy(seq["rot6d_init"].copy())
    tr = torch.from_numpy(seq["T_init"].copy())
    mom = E.mesh_moments(verts)
    hist = []
    for it in range(3):
        st = E.smooth_terms(rot.numpy(), tr.numpy(), 1.0, mom, len(verts), B, 10.0)
        g6 = E.rot6d_backward(rot.numpy(), st[:, 3:12])
        rot = rot - 0.05 * torch.from_numpy(g6).float()
        tr = tr - 0.05 * torch.from_numpy(st[:, 0:3]).float().reshape(-1, 1, 3)
        hist.append(float(st[:, 13].sum()))
    return rot.numpy(), tr.numpy(), hist


@pytest.mark.parametrize("B", [7, 10])
def test_two_rank_halo_exchange_equals_single_shard(B):
    E.lib()  # build the emu library before forking
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    rot2, tr2, hist2, keep = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rot1, tr1, hist1 = _single(B)
    assert keep == 201.0
    assert np.array_equal(rot1, rot2) and np.array_equal(tr1, tr2)
    assert np.allclose(hist1, hist2, rtol=1e-12)


# ------------------------------------------------------------------------------------------ cost-weighted ranges
def test_explicit_bounds_and_balanced_partition():
    from dynhor_b200.sharding import balanced_bounds, frame_costs_from_blocks
    sh = FrameShard(1, 3, 10, bounds=[0, 2, 7, 10])
    assert (sh.start, sh.stop, sh.B) == (2, 7, 5) and sh.of_rank(2).B == 3
    for bad in ([0, 2, 2, 10], [0, 5, 10], [1, 2, 7, 10], [0, 2, 7, 11]):
        with pytest.raises(ValueError):
            FrameShard(0, 3, 10, bounds=bad)
    # uniform cost -> the by-count split
    assert balanced_bounds(np.ones(4096), 8) == [512 * r for r in range(9)]
    # frames twice as expensive in the second half: ranges there are half as long
    cost = np.concatenate([np.ones(600), 2 * np.ones(600)])
    b = balanced_bounds(cost, 6)
    loads = [cost[b[r]:b[r + 1]].sum() for r in range(6)]
    assert max(loads) - min(loads) <= 2.0 and b[0] == 0 and b[-1] == 1200
    assert b[1] - b[0] == 300 and b[-1] - b[-2] == 150
    # degenerate costs still give a valid partition with >= 1 frame per rank
    for c in (np.zeros(5), np.array([0, 0, 0, 0, 9.0]), np.array([np.nan, 1, 1, 1, 1.0])):
        bb = balanced_bounds(c, 5)
        assert bb == [0, 1, 2, 3, 4, 5]
    rng = np.random.default_rng(0)
    for _ in range(50):
        n, w = int(rng.integers(8, 400)), int(rng.integers(1, 9))
        bb = balanced_bounds(rng.random(n) + 0.01, w)
        assert len(bb) == w + 1 and bb[0] == 0 and bb[-1] == n and all(y > x for x, y in zip(bb, bb[1:]))
    # probe blocks -> per-frame costs: block k covers frames B*k/n .. B*(k+1)/n of the range, extra spread evenly
    c = frame_costs_from_blocks([1.0, 3.0], 10, 17, extra_ms=0.7)
    assert len(c) == 7 and np.isclose(c.sum(), 4.7) and np.allclose(c[:3], 1 / 3 + 0.1) and np.allclose(c[3:], 0.75 + 0.1)


def test_fx128_exact_sum_is_order_and_partition_independent():
    """The scale gradient is summed exactly (dh_core.h Fx128): any grouping of the per-frame terms -- one GPU, or
    per-rank partial sums added in rank order -- gives the same bits.  Host mirror == C arithmetic."""
    from dynhor_b200.sharding import fx128_from_float, fx128_sum
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.normal(size=500) * 10.0 ** rng.integers(-12, 6, size=500), [0.0, -0.0, 1e-30, -1e-30,
                        2.0 ** -70, -(2.0 ** -70), 123456789.125, -0.5]])
    hi, lo, val = E.fx_sum(x)
    assert fx128_sum([fx128_from_float(float(v)) for v in x]) == val
    for perm in (x[::-1], rng.permutation(x)):
        assert E.fx_sum(perm) == (hi, lo, val)
    for cuts in ([0, 100, 508], [0, 1, 2, 3, 400, 508], [0, 254, 508]):
        parts = [E.fx_sum(x[a:b])[:2] for a, b in zip(cuts, cuts[1:])]
        assert fx128_sum(parts) == val
    # close to the correctly rounded double sum
    import math
    assert abs(val - math.fsum(x)) <= 1e-15 * max(1.0, abs(math.fsum(x))) + 508 * 2.0 ** -64


def _scale_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dynhor_b200.sharding import allgather_equal, fx128_sum
        rng = np.random.default_rng(3)
        g = rng.normal(size=37) * 1e-3                      # per-frame scale gradients of the whole sequence
        sh = FrameShard(rank, world, 37, bounds=[0, 11, 37])
        hi, lo, _ = E.fx_sum(g[sh.start:sh.stop])
        part = torch.tensor([hi, lo], dtype=torch.uint64).view(torch.int64)
        parts = allgather_equal(part, sh).numpy().view(np.uint64)   # what DH_SCALE_DEFERRED hands to dh_scale_apply
        total = fx128_sum([(int(a), int(b)) for a, b in parts])
        q.put((rank, total))
    finally:
        dist.destroy_process_group()


def test_two_rank_scale_gradient_equals_single_shard():
    E.lib()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_scale_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = np.random.default_rng(3).normal(size=37) * 1e-3
    single = E.fx_sum(g)[2]
    assert got[0] == got[1] == single


def test_partition_policy():
    """balance="auto": the cost probe from 64 iterations on (it costs ~5 iterations and buys ~10 % of each); "probe" /
    "count" force the choice; one rank never probes; anything else is an error (before any work is done)."""
    import pytest
    from dynhor_b200.sharding import PROBE_MIN_ITERATIONS, wants_probe
    assert PROBE_MIN_ITERATIONS == 64
    assert not wants_probe("auto", 20, 8) and wants_probe("auto", 64, 8) and wants_probe("auto", 200, 2)
    assert wants_probe("probe", 1, 2) and not wants_probe("count", 10 ** 6, 8)
    for b in ("auto", "probe", "count"):
        assert not wants_probe(b, 1000, 1)
    with pytest.raises(ValueError):
        wants_probe("cost", 100, 8)
