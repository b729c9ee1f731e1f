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
