"""2-GPU test (skipped on a single-GPU box): frame-sharded joint optimisation over NCCL with the peer-to-peer
mailboxes and with the host-driven NCCL exchange must both reproduce the single-GPU run bit for bit -- poses, loss
history and, with optimize_object_scale, the shared scale -- for ranges cut by count and by the cost probe."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

LW = {"lw_sil_obj": 1.0, "lw_smooth_obj": 10.0}
ITERS = 10


def _gpu_render_fn(vc, faces, K, size):
    from dynhor_b200.renderer import Renderer
    B = len(vc)
    r = Renderer(image_size=size, K=torch.from_numpy(K).cuda(), R=torch.eye(3)[None].cuda(),
                 t=torch.zeros(1, 3).cuda(), orig_size=1, anti_aliasing=False)
    with torch.no_grad():
        return r(torch.from_numpy(vc).cuda(), torch.from_numpy(faces).cuda()[None].repeat(B, 1, 1),
                 mode="silhouettes").cpu().numpy()


def _worker(rank, world, port, seq, halo, q, scale_opt=False, balance="count", iters=10, no_p2p=False):
    import torch.distributed as dist
    from dynhor_b200 import synth
    from dynhor_b200.jointopt import joint_optimize
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if no_p2p:
        os.environ["DH_FORCE_NO_P2P"] = "1"   # pretend the devices cannot map each other's memory
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        params = synth.to_object_parameters(seq)
        B = len(params)
        model, evo = joint_optimize(params, objvertices=seq["verts"], objfaces=np.stack([seq["faces"]] * B),
                                    loss_weights=LW, num_iterations=iters, lr=1e-3 if scale_opt else 1e-4, board=None,
                                    halo=halo, optimize_object_scale=scale_opt, balance=balance,
                                    _rebalance_to=[0, 7, len(params)] if iters >= 64 else None)
        if rank == 0:
            q.put((model.rotations_object.detach().cpu().numpy(), model.translations_object.detach().cpu().numpy(),
                   evo, float(model.int_scales_object.detach()), model.frame_shard.bounds))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("halo,scale_opt,balance,iters", [("p2p", False, "count", 10), ("nccl", False, "count", 10),
                                                           ("p2p", True, "probe", 10), ("nccl", True, "probe", 10),
                                                           ("p2p", True, "probe", 70), ("p2p-unavailable", True, "probe", 10)])
def test_two_gpu_sharded_equals_single_gpu(halo, scale_opt, balance, iters):
    """iters = 70: long enough for the mid-run re-partition (after 16 iterations the ranges are re-cut from the ranks'
    own clocks and frames change hands together with their Adam moments) -- still the single-GPU bits."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from dynhor_b200 import synth
    from dynhor_b200.jointopt import joint_optimize
    B = 21  # ragged split 11 + 10
    seq = synth.make_sequence(B, mesh="ico3", seed=4, render_fn=_gpu_render_fn, size=128, period=60)
    params = synth.to_object_parameters(seq)
    model, evo1 = joint_optimize(params, objvertices=seq["verts"], objfaces=np.stack([seq["faces"]] * B),
                                 loss_weights=LW, num_iterations=iters, lr=1e-3 if scale_opt else 1e-4, board=None,
                                 optimize_object_scale=scale_opt)
    scale1 = float(model.int_scales_object.detach())
    rot1 = model.rotations_object.detach().cpu().numpy()
    tr1 = model.translations_object.detach().cpu().numpy()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    no_p2p = halo == "p2p-unavailable"    # halo="p2p" requested, no peer access: every rank falls back to the NCCL exchange
    halo = "p2p" if no_p2p else halo
    procs = [ctx.Process(target=_worker, args=(r, 2, port, seq, halo, q, scale_opt, balance, iters, no_p2p))
             for r in range(2)]
    for p in procs:
        p.start()
    rot2, tr2, evo2, scale2, bounds = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert rot2.shape == rot1.shape
    assert np.array_equal(rot1, rot2) and np.array_equal(tr1, tr2)
    assert scale1 == scale2 and (scale1 != 1.0) == scale_opt          # the shared scale: same bits on every rank
    assert bounds[0] == 0 and bounds[-1] == B and len(bounds) == 3 and (iters < 64 or bounds[1] == 7)
    assert np.allclose(evo1["loss"], evo2["loss"], rtol=1e-12)
    assert np.allclose(evo1["iou_object"], evo2["iou_object"], rtol=1e-12)


def _dino_worker(rank, world, port, q):
    import torch.distributed as dist
    from dynhor_b200 import synth
    from dynhor_b200.dino_match import build_bank, dino_cos_topk, dino_topk_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        N, Fm, P, D, k = 300, 20, 24, 64, 5
        d = synth.make_dino_features(N, Fm, P, D, seed=7, device="cuda")
        fb = build_bank(d["frames"], d["masks"])
        sizes = [170, 130]                                   # ragged template slices
        a = sum(sizes[:rank])
        tb_local = build_bank(d["templ"][a:a + sizes[rank]])
        vals, idx = dino_topk_sharded(fb, tb_local, k, rank, world, sizes)
        _, v_all, i_all = dino_cos_topk(fb, build_bank(d["templ"]), k)
        if rank == 0:
            q.put((bool(torch.equal(idx, i_all)), float((vals - v_all).abs().max())))
    finally:
        dist.destroy_process_group()


def test_two_gpu_template_sharded_topk_equals_whole_bank():
    """The reference-sized template bank does not have to live on one GPU: templates sharded over ranks, per-rank
    top-k, one all_gather of the [Fm,k] lists, merge_topk == the top-k of the whole bank."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [ctx.Process(target=_dino_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same_idx, dv = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert same_idx and dv < 1e-5
