#!/usr/bin/env python
"""bench.py -- joint pose-optimisation throughput (frame-iterations / second) on N B200s of one box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames-per-gpu F]

A "step" is one fused optimisation iteration (forward + backward + Adam, jointopt.py:144-160) over every frame
of the sequence.  Workload at N=1 = BASELINE.json configs[1]: custom_shoes-shaped joint optimisation, 300 frames
of a 480x640 sequence, 5k-vertex mesh (V=5002, F=10000), 256x256 ROIs rendered at 512x512 with anti-aliasing,
loss weights of configs/custom_shoes.yaml.  N>1: weak scaling, 300 frames per GPU, frame-range sharding with a
one-frame pose halo exchange per iteration (no other data-path collective).

One JSON line on stdout (rank 0).  `--impl reference` times the CPU oracle (the restatement of the reference's
CPU-incapable path) on the host cores over a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

LW = {"lw_sil_obj": 1.0, "lw_smooth_obj": 10.0}   # configs/custom_shoes.yaml:17-18
LW_CORR = 0.01                                    # builder-defined correspondence term (dynhor_b200/corr.py)
LR = 1e-4                                         # configs/custom_shoes.yaml:15
MESH = "uv50x100"                                 # V=5002, F=10000 ("5k-vertex mesh")
H, W, S = 480, 640, 256
METRIC = "jointopt frame-iters/sec (fwd+bwd+Adam)"
UNIT = "frame-iters/s"


def algorithmic_bytes_per_frame(V, S=256, C=0):
    """SURVEY.md 8(d) kernel-boundary model, per frame-iteration, split by kernel (fp32, u8 coverage, i32 index);
    + 24 bytes per correspondence record."""
    SS = 2 * S
    return {
        "corr": 24 * C,
        "project": 12 * V,                          # writes projected vertices
        "raster": 12 * V + 5 * SS * SS + SS * SS + 8 * S * S,  # raster fwd + the loss kernel fused into it
        "backward": 5 * SS * SS + 8 * S * S + 12 * V + 12 * V,
        "pose_update": 12 * V + 288,
        "total": 60 * V + 11 * SS * SS + 16 * S * S + 288 + 24 * C,
    }


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.first = index, [], None, 0

    def mark(self):
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        rows = self.rows[self.first:] if len(self.rows) - self.first >= 2 else self.rows[-3:]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- data
def gpu_render_fn(vc, faces, K, size):
    from dynhor_b200.renderer import Renderer
    B = len(vc)
    r = Renderer(image_size=size, K=torch.from_numpy(K).cuda(), R=torch.eye(3)[None].cuda(),
                 t=torch.zeros(1, 3).cuda(), orig_size=1, anti_aliasing=False)
    with torch.no_grad():
        return r(torch.from_numpy(vc).cuda(), torch.from_numpy(faces).cuda()[None].repeat(B, 1, 1),
                 mode="silhouettes").cpu().numpy()


def oracle_render_fn(vc, faces, K, size):
    from oracle import nr_oracle
    B = len(vc)
    r = nr_oracle.Renderer(image_size=size, K=torch.from_numpy(K), R=torch.eye(3)[None], t=torch.zeros(1, 3),
                           orig_size=1, anti_aliasing=False)
    return r(torch.from_numpy(vc), torch.from_numpy(faces)[None].repeat(B, 1, 1), mode="silhouettes").numpy()


def loss_weights(C):
    return dict(LW, lw_corr_obj=LW_CORR) if C > 0 else dict(LW)


def build_model(seq):
    from dynhor_b200 import synth
    from dynhor_b200.jointopt import Joint_Optimizer
    params = synth.to_object_parameters(seq)
    B = len(params)
    corr = torch.from_numpy(seq["correspondences"]) if "correspondences" in seq else None
    return Joint_Optimizer(
        correspondences=corr,
        translations_object=torch.cat([p["translations"] for p in params]),
        rotations_object=torch.cat([p["rotations"] for p in params]),
        verts_object_og=torch.from_numpy(seq["verts"]),
        faces_object=torch.from_numpy(seq["faces"]),
        camintr_rois_object=torch.cat([p["K_roi"][:, 0] for p in params]),
        target_masks_object=torch.cat([p["target_masks"] for p in params]),
        int_scale_init=1, optimize_object_scale=False)


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_baseline(frames=None, iters=1, C=0):
    """The CPU oracle (kind "port": the reference cannot run on CPU) on a bounded sample of
    the same workload: `frames` custom_shoes-shaped frames x `iters` full iterations, all host threads."""
    from dynhor_b200 import synth
    from oracle import jointopt_oracle as jo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    frames = frames or max(2, min(cores, 64))
    seq = synth.make_sequence(frames, H, W, mesh=MESH, seed=0, render_fn=oracle_render_fn, period=300)
    corr = synth.make_correspondences(seq, C, seed=0) if C > 0 else None
    orc = jo.JointOptOracle(seq["rot6d_init"], seq["T_init"], seq["verts"], seq["faces"], seq["K_roi"],
                            seq["target_masks"], lr=LR, correspondences=corr)
    lw = loss_weights(C)
    t0 = time.perf_counter()
    for _ in range(iters):
        orc.step(lw)
    dt = time.perf_counter() - t0
    return {"value": frames * iters / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{frames} frames x {iters} iteration(s) of the N=1 workload (5k-vertex mesh, 512x512 AA "
                      f"raster, {C} correspondences/frame), {dt:.1f} s of CPU work, OpenMP + torch threads = {cores}"}, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    cores = os.cpu_count() or 1
    frames = max(2, min(cores // 2, 32))
    cb_rows = []
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(frames=2, iters=1, C=args.corr)
    t_total = 0.0
    for _ in range(steps):
        cb, dt = cpu_baseline(frames=frames, iters=1, C=args.corr)
        cb_rows.append(cb)
        t_total += dt
    value = frames * steps / t_total
    cb = dict(cb_rows[-1])
    cb["value"] = value
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1000.0 * t_total / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "custom_shoes-shaped joint optimisation (BASELINE configs[1] frame shape): "
                               f"{frames}-frame bounded sample per step, 480x640, 5k-vertex mesh, 256x256 ROI at "
                               f"512x512 AA, {args.corr} correspondences per frame; CPU oracle port of the reference "
                               "path (reference itself is CUDA-only)",
                   "frames_per_step": frames, "correspondences_per_frame": args.corr},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch.distributed as dist
    from dynhor_b200 import synth
    from dynhor_b200.jointopt import FusedJointOpt, joint_optimize
    from dynhor_b200.sharding import FrameShard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    Bl = args.frames_per_gpu
    B_total = Bl * world
    shard = FrameShard(rank, world, B_total)
    period, offset = B_total, shard.start
    if args.emulate_shard and world == 1:
        # diagnostic: the frames rank r of a w-GPU run would own, timed stand-alone on one GPU (no halo)
        r, w = (int(v) for v in args.emulate_shard.split("/"))
        period, offset = Bl * w, Bl * r
    # weak scaling: the object's motion repeats every `frames_per_gpu` frames, so every rank's range holds one
    # full period of it (same amount of work per GPU at every N; --traj-period total: one period over the whole
    # sequence, where the ranks' frame ranges differ in content and the slowest one sets the pace)
    traj = Bl if args.traj_period == "per-gpu" else period
    seq = synth.make_sequence(Bl, H, W, mesh=MESH, seed=0, render_fn=gpu_render_fn, period=period,
                              frame_offset=offset, traj_period=traj)
    V, F = len(seq["verts"]), len(seq["faces"])
    C = args.corr
    lw = loss_weights(C)
    if C > 0:
        seq["correspondences"] = synth.make_correspondences(seq, C, seed=shard.start)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput: K fused iterations, inputs already in HBM
    model = build_model(seq)
    fused = FusedJointOpt(model, lw, LR, args.steps + args.warmup + 64, shard=shard, halo=args.halo)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.7)          # nvidia-smi needs a moment before its first sample
    fused.run(args.warmup, use_graph=True)
    barrier()
    if rank == 0:
        sampler.mark()           # samples from here on are inside the timed region
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    fused.run(args.steps, use_graph=True)
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    value = B_total * args.steps / (ms / 1000.0)
    hist = fused.history()

    # ---- per-kernel times (CUDA events on the launch stream) -> roofline of the dominant kernel
    prof = fused.profile(5)
    fused.release()
    ab = algorithmic_bytes_per_frame(V, S, C)
    top = max(("project", "raster", "backward", "pose_update", "corr"), key=lambda k: prof[k])
    peak, peak_kind = measured_peaks()
    achieved = ab[top] * Bl / (prof[top] / 1000.0) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            t = json.load(open(tpath)).get(top)
            traffic = t["dram_bytes_per_frame"] * Bl if t else None
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_" + top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": f"{peak_kind} hbm_gbs",
                "algorithmic_bytes_per_launch": ab[top] * Bl, "kernel_ms": prof[top],
                "kernel_ms_all": prof,
                "kernel_gbs_all": {k: ab[k] * Bl / (prof[k] / 1000.0) / 1e9
                                   for k in ("project", "raster", "backward", "pose_update", "corr") if prof[k] > 0},
                "whole_step": {"algorithmic_bytes": ab["total"] * Bl,
                               "achieved": ab["total"] * Bl / (ms / args.steps / 1000.0) / 1e9,
                               "frac": ab["total"] * Bl / (ms / args.steps / 1000.0) / 1e9 / peak}}

    # ---- end to end through the public call with HOST buffers (H2D of inputs, D2H of results inside the timing)
    full = seq if world == 1 else None
    if world > 1:
        # every rank holds the full-sequence host inputs, like run.py would; masks of other ranks are rebuilt
        # from the all-gathered local masks
        m_local = torch.from_numpy(seq["target_masks"]).cuda()
        ms_all = [torch.empty_like(m_local) for _ in range(world)]
        dist.all_gather(ms_all, m_local)
        full = synth.make_sequence(B_total, H, W, mesh=MESH, seed=0, render_fn=None, period=B_total,
                                   traj_period=traj)
        full["target_masks"] = torch.cat(ms_all).cpu().numpy()
        if C > 0:  # only this rank's frames are read by joint_optimize; the others are placeholders
            full["correspondences"] = np.zeros((B_total, C, 6), np.float32)
            full["correspondences"][shard.start:shard.stop] = seq["correspondences"]
    params = synth.to_object_parameters(full)
    keys = ("rotations", "translations", "K_roi", "target_masks") + (("correspondences",) if C > 0 else ())
    for p in params:
        for k in keys:
            p[k] = p[k].pin_memory()
    faces_b = np.stack([full["faces"]] * B_total)
    e2e_iters = args.steps
    h2d = sum(p[k].numel() * p[k].element_size() for p in params[shard.start:shard.stop] for k in keys)
    h2d += full["verts"].nbytes + faces_b.nbytes
    joint_optimize(params, objvertices=full["verts"], objfaces=faces_b, loss_weights=lw, num_iterations=2, lr=LR)
    barrier()
    t0 = time.perf_counter()
    model2, evo = joint_optimize(params, objvertices=full["verts"], objfaces=faces_b, loss_weights=lw,
                                 num_iterations=e2e_iters, lr=LR, board=None)
    rot_h = model2.rotations_object.detach().cpu()
    tr_h = model2.translations_object.detach().cpu()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    d2h = rot_h.numel() * 4 + tr_h.numel() * 4 + 4 * 8 * e2e_iters
    dt_t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
    dt = float(dt_t.item())
    e2e = {"value": B_total * e2e_iters / dt, "unit": UNIT, "h2d_bytes_per_step": h2d / e2e_iters,
           "d2h_bytes_per_step": d2h / e2e_iters, "iterations": e2e_iters, "seconds": dt,
           "final_loss": evo["loss"][-1]}

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb, _ = cpu_baseline(C=C)
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"custom_shoes-shaped joint pose optimisation, {Bl} frames per GPU "
                                   f"({B_total} total) {H}x{W}, {MESH} mesh (V={V}, F={F}), 256x256 ROI rendered "
                                   f"512x512 + 2x2 pool, {C} correspondences per frame, lw_sil 1 / lw_smooth 10"
                                   + (f" / lw_corr {LW_CORR}" if C > 0 else "") + ", lr 1e-4 (BASELINE configs[1])",
                       "frames_per_gpu": Bl, "frames_total": B_total, "correspondences_per_frame": C,
                       "parallelism": f"frame-shard x{world}",
                       "trajectory_period_frames": traj,
                       "halo": fused.halo_mode,
                       "l2": "per-step working set (face-index maps 1 MB/frame + bins) exceeds the 126 MB L2; "
                             "no explicit flush", "cuda_graph": True},
            "roofline": roofline, "e2e": e2e, "clocks": clocks,
            "gpu_launches": (9 + (1 if C > 0 else 0)) * args.steps,
            "loss_first_last": [hist["loss"][0], hist["loss"][-1]],
            "iou_first_last": [hist["iou_object"][0], hist["iou_object"][-1]],
        }
        if cb is not None:
            out["cpu_baseline"] = cb
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_dino(args):
    """Secondary workload (BASELINE configs[3]): DINO ViT-S/14 patch-feature matching, 1k templates x 300 frames,
    bf16 similarity GEMM (K = 1369*384) + fused top-10.  One JSON line; `value` = template-frame pairs / second."""
    from dynhor_b200 import synth
    from dynhor_b200.dino_match import build_bank, dino_cos_topk
    N, Fm, P, D, k = 1000, 300, 1369, 384, 10
    torch.cuda.set_device(0)
    d = synth.make_dino_features(N, Fm, P, D, seed=0, device="cuda")
    tb = build_bank(d["templ"])
    fb = build_bank(d["frames"], d["masks"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(max(args.warmup, 3)):
        dino_cos_topk(fb, tb, k)
    times = []
    for _ in range(args.steps):
        flush.zero_()                      # 256 MB write: evicts the banks from the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s, v, i = dino_cos_topk(fb, tb, k)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    flops = 2.0 * N * Fm * P * D
    bytes_ = 2.0 * P * D * (N + Fm)
    hbm, kind = measured_peaks()
    tf_peak = 1606.8
    try:
        tf_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        pass
    ok = bool(torch.equal(i[:, 0].cpu(), d["match"]))
    import ctypes
    from dynhor_b200 import _lib
    plan = (ctypes.c_int32 * 8)()
    _lib.load().dh_dino_plan_info(N, Fm, P * D, plan)
    # CPU baseline: the reference expression verbatim on a bounded sample of frames
    from oracle import dino_oracle
    nf = 2
    t0 = time.perf_counter()
    dino_oracle.dino_cos_topk(d["frames"][:nf].cpu(), d["masks"][:nf].cpu(), d["templ"].cpu(), k)
    cpu_s = (time.perf_counter() - t0) / nf
    print(json.dumps({
        "metric": "DINO template-frame pairs scored / second (bf16 GEMM + fused top-k)", "value": N * Fm / (ms / 1e3),
        "unit": "pairs/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
        "higher_is_better": True, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "DINO ViT-S/14 patch-feature pose initialisation: 1000 templates x 300 frames, "
                               "P=1369, D=384, top-10 (BASELINE configs[3])", "l2": "256 MB flush between steps"},
        "roofline": {"bound": "hbm", "achieved": bytes_ / (ms / 1e3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": bytes_ / (ms / 1e3) / 1e9 / hbm, "peak_source": f"{kind} hbm_gbs",
                     "tensor": {"achieved_tflops": flops / (ms / 1e3) / 1e12, "peak_tflops": tf_peak,
                                "frac": flops / (ms / 1e3) / 1e12 / tf_peak}},
        "cpu_baseline": {"value": N / cpu_s, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "reference",
                         "sample": f"{nf} frames x 1000 templates, pose_initializtion.py:295-296 verbatim + topk"},
        "plan": dict(zip(["m_tiles", "n_pairs", "k_slices", "kblocks", "kb_per_slice", "cluster", "ctas_sized_for",
                          "ldc"], list(plan))),
        "planted_match_rank0": ok, "gpu_launches": 3 * args.steps}))


def run_preprocess(args):
    """Secondary workload: run.py:26-72 process_input for a whole sequence -- tight boxes,
    ROIAlign crops of object mask / hand mask / image, tri-state target masks.  One JSON line; `value` = frames/s with
    the inputs resident in HBM, `e2e` = from host arrays (H2D of masks + images, D2H of every output)."""
    from dynhor_b200.preprocess import process_input, process_input_batched
    from oracle import roi_oracle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import roi_scenes
    B = args.frames_per_gpu
    torch.cuda.set_device(0)
    base = roi_scenes(12, H, W, 0)
    images = [base[0][i % 12] for i in range(B)]
    objs = [base[1][i % 12] for i in range(B)]
    hands = [base[2][i % 12] for i in range(B)]
    im_d = torch.from_numpy(np.stack(images)).cuda()
    ob_d = torch.from_numpy(np.stack([o == 255 for o in objs])).cuda().to(torch.uint8)   # bit masks resident in HBM
    hb_d = torch.from_numpy(np.stack([h == 255 for h in hands])).cuda().to(torch.uint8)
    for _ in range(max(args.warmup, 3)):
        process_input_batched(im_d, ob_d, hb_d, masks_are_bits=True)
    steps = max(1, min(args.steps, 50))
    times = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = process_input_batched(im_d, ob_d, hb_d, masks_are_bits=True)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    t0 = time.perf_counter()
    out = process_input(images, objs, hands)
    e2e_s = time.perf_counter() - t0
    # bytes: every input byte once (masks as float64 arrays like run.py's load_data + uint8 image), every output once
    in_bytes = B * H * W * (1 + 1 + 3)          # device-resident: uint8 bit masks + uint8 image
    h2d_bytes = in_bytes                           # the host wrapper compares the float64 masks into pinned uint8
    out_bytes = B * (256 * 256 * (1 + 4 + 1 + 12) + 32)
    hbm, kind = measured_peaks()
    nf = min(B, 16)
    t0 = time.perf_counter()
    ref = roi_oracle.process_input(images[:nf], objs[:nf], hands[:nf])
    cpu_s = (time.perf_counter() - t0) / nf
    ok = all(np.array_equal(out[i]["target_crop_mask"], ref[i]["target_crop_mask"]) for i in range(nf))
    print(json.dumps({
        "metric": "ROI preprocessing frames / second (boxes + ROIAlign crops + target masks)", "value": B / (ms / 1e3),
        "unit": "frames/s", "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
        "higher_is_better": True, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"run.py process_input: {B} frames {H}x{W}, object + hand masks + RGB image -> 256x256 "
                               "crop mask, crop image and tri-state target mask", "l2": "inputs (1.5 MB/frame, 461 MB) exceed L2"},
        "roofline": {"bound": "hbm", "achieved": (in_bytes + out_bytes) / (ms / 1e3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": (in_bytes + out_bytes) / (ms / 1e3) / 1e9 / hbm, "peak_source": f"{kind} hbm_gbs",
                     "traffic": None, "algorithmic_bytes_per_launch": in_bytes + out_bytes},
        "e2e": {"value": B / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": out_bytes},
        "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "frames/s", "cores": os.cpu_count(), "kind": "reference",
                         "sample": f"{nf} frames, run.py:26-72 restated line by line over torchvision's CPU roi_align "
                                   "(what detectron2's ROIAlign executes)"},
        "matches_oracle": bool(ok), "gpu_launches": 3 * steps}))


def _quiet_stdout():
    """Route everything libraries print on fd 1 (NCCL's version banner, tqdm) to stderr; the JSON line is written
    to the real stdout at the end, so that stdout carries exactly one line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real, "w")
    import builtins
    _print = builtins.print

    def emit(*a, **k):
        if k.get("file") is None:
            k["file"] = out
            k["flush"] = True
        _print(*a, **k)
    return emit


def main():
    global print, MESH, H, W
    print = _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)     # configs/custom_shoes.yaml:14 joint_num_iterations
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-gpu", type=int, default=300)
    ap.add_argument("--corr", type=int, default=10000,
                    help="correspondences per frame (builder-defined reprojection term; 0 = the reference's two terms)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mesh", default=MESH, choices=["uv50x100", "uv100x200"],
                    help="uv100x200 = the 20k-vertex mesh of BASELINE configs[2]")
    ap.add_argument("--camera", default=f"{H}x{W}", help="full-frame camera HxW (configs[2]: 1080x1920)")
    ap.add_argument("--traj-period", default="per-gpu", choices=["per-gpu", "total"],
                    help="frames after which the synthetic motion repeats: frames-per-gpu (default) or the whole sequence")
    ap.add_argument("--emulate-shard", default="", help="r/w: time the frames of rank r of a w-GPU run on one GPU")
    ap.add_argument("--workload", default="jointopt", choices=["jointopt", "dino", "preprocess"])
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"])
    args = ap.parse_args()
    MESH = args.mesh
    H, W = (int(v) for v in args.camera.split("x"))
    if args.workload == "dino":
        run_dino(args)
    elif args.workload == "preprocess":
        run_preprocess(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        run_ours(args)


if __name__ == "__main__":
    main()
