#!/usr/bin/env python
"""bench.py -- joint pose-optimisation throughput (frame-iterations / second) on N B200s of one box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling auto|weak|strong]
                    [--workload jointopt|stage1|dino|preprocess]

A "step" is one fused optimisation iteration (forward + backward + Adam, jointopt.py:144-160) over every frame
of the sequence.
  N = 1   BASELINE.json configs[1]: custom_shoes-shaped joint optimisation, 300 frames of a 480x640 sequence, 5k-vertex
          mesh (V=5002, F=10000), 256x256 ROIs rendered at 512x512 with anti-aliasing, loss weights of
          configs/custom_shoes.yaml, 10k correspondences per frame.  The line also carries the DINO matcher's numbers
          (configs[3]) under "secondary".
  N > 1   BASELINE.json configs[4], STRONG scaling: one 4096-frame sequence with 50k correspondences per frame (one period
          of the synthetic motion over the whole sequence, so the ranks' frame ranges differ in content), cut into
          contiguous frame ranges of equal measured cost (one probe pass), one-frame pose halo per iteration written
          peer-to-peer over NVLink.  After the timed region the line reports `shard_equals_single` (3 iterations on all
          ranks compared bit for bit with the whole sequence on rank 0 alone) and `single_gpu_same_config` (that run's
          throughput: the denominator of the strong-scaling efficiency).  `--scaling weak` keeps 300 frames per GPU.

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
    if args.corr is None:
        args.corr = 10000
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
def make_range(start, stop, B_total, C, with_masks=True):
    """Frames [start, stop) of the B_total-frame synthetic sequence (one period of the motion over the WHOLE
    sequence): every frame's data depends on (seed, global frame number) only, so any rank -- or rank 0 alone, for
    the sharded-equals-single check -- builds the same frames."""
    from dynhor_b200 import synth
    seq = synth.make_sequence(stop - start, H, W, mesh=MESH, seed=0, render_fn=gpu_render_fn if with_masks else None,
                              period=B_total, frame_offset=start, traj_period=B_total)
    if C > 0:
        seq["correspondences"] = synth.make_correspondences(seq, C, seed=0, frame_offset=start, device="cuda")
    return seq


def model_from(seq):
    from dynhor_b200.jointopt import Joint_Optimizer
    corr = seq.get("correspondences")
    if corr is not None and not torch.is_tensor(corr):
        corr = torch.from_numpy(corr)
    return Joint_Optimizer(
        correspondences=corr,
        translations_object=torch.from_numpy(seq["T_init"]),
        rotations_object=torch.from_numpy(seq["R_init"]),
        verts_object_og=torch.from_numpy(seq["verts"]),
        faces_object=torch.from_numpy(seq["faces"]),
        camintr_rois_object=torch.from_numpy(seq["K_roi"]),
        target_masks_object=torch.from_numpy(seq["target_masks"]),
        int_scale_init=1, optimize_object_scale=False)


def host_parameters(seq, start, B_total, C):
    """The per-frame dict list joint_optimize takes (pose_initializtion.py:460-471) for the whole sequence, with
    pinned host tensors for the frames of `seq` (global numbers start ...) and references to one of them elsewhere
    (frames another rank owns are never read by this rank)."""
    from dynhor_b200 import synth
    if C > 0 and torch.is_tensor(seq["correspondences"]):
        seq = dict(seq, correspondences=seq["correspondences"].cpu().numpy())
    local = synth.to_object_parameters(seq)
    keys = ("rotations", "translations", "K_roi", "target_masks") + (("correspondences",) if C > 0 else ())
    for p in local:
        for k in keys:
            p[k] = p[k].pin_memory()
    full = [local[0]] * start + local + [local[-1]] * (B_total - start - len(local))
    nbytes = lambda fr: sum(p[k].numel() * p[k].element_size() for p in fr for k in keys)  # noqa: E731
    return full, nbytes


def run_ours(args):
    import torch.distributed as dist
    from dynhor_b200.jointopt import FusedJointOpt, joint_optimize
    from dynhor_b200.sharding import (FrameShard, allgather_equal, balanced_bounds, frame_costs_from_blocks,
                                       rescale_costs)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    scaling = args.scaling if args.scaling != "auto" else ("strong" if world > 1 else "single")
    if scaling == "strong":          # BASELINE configs[4]: fixed total, 50k correspondences per frame pair
        B_total = args.frames_total or 4096
        C = 50000 if args.corr is None else args.corr
        workload = "scaling sweep (BASELINE configs[4])"
    else:                            # BASELINE configs[1] on one GPU; weak: the same number of frames per GPU
        B_total = args.frames_per_gpu * world
        C = 10000 if args.corr is None else args.corr
        workload = "custom_shoes-shaped joint pose optimisation (BASELINE configs[1])"
    lw = loss_weights(C)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- frame ranges: by count on one GPU; by measured cost across GPUs (one probe pass over the equal-count
    #      ranges: the ranks' frames differ in content and every rank steps at the pace of the slowest range)
    shard = FrameShard(rank, world, B_total)
    if args.emulate_shard and world == 1:
        # diagnostic: the frames rank r of a w-GPU run would own (by count), timed stand-alone on one GPU
        r, w = (int(v) for v in args.emulate_shard.split("/"))
        B_total = args.frames_total or args.frames_per_gpu * w
        es = FrameShard(r, w, B_total)
        seq = make_range(es.start, es.stop, B_total, C)
        shard = FrameShard(0, 1, es.B)
    else:
        seq = None
    probe_ms, calib_ms = None, None
    if world > 1 and args.balance == "probe":
        seq0 = make_range(shard.start, shard.stop, B_total, 0)
        nblocks = max(1, min(16, (B_total // world) // 32))
        with FusedJointOpt(model_from(seq0), dict(LW), LR, 0, shard=shard, keep_sum=1.0, exchange=False) as pr:
            ms_blocks = torch.from_numpy(pr.probe(nblocks)).cuda()
        ms_all = allgather_equal(ms_blocks, shard).cpu().numpy()
        cost = np.concatenate([frame_costs_from_blocks(ms_all[r, :nblocks], shard.bounds[r], shard.bounds[r + 1])
                               for r in range(world)])
        probe_ms = [float(ms_all[r, :nblocks].sum()) for r in range(world)]
        shard = shard.with_bounds(balanced_bounds(cost, world))
        del seq0, pr
        # second cut, like joint_optimize's: 16 iterations on the probed ranges, every rank's own time per iteration
        # (device clock, waits excluded) rescales the cost profile; the calibration run is thrown away
        seq1 = make_range(shard.start, shard.stop, B_total, C)
        with FusedJointOpt(model_from(seq1), lw, LR, 16, shard=shard, halo=args.halo) as cal:
            cal.run(16)
            t = torch.tensor([float(np.mean(cal.compute_ms(8, 16)))], dtype=torch.float64, device="cuda")
            cal.check_status()
        t_all = allgather_equal(t, shard).reshape(-1).cpu().numpy()
        calib_ms = [float(v) for v in t_all]
        if t_all.max() > 1.01 * t_all.mean():      # (8-iteration means of a device clock: the noise is well below 1 %)
            nb = balanced_bounds(rescale_costs(cost, shard.bounds, t_all), world)
            if nb != shard.bounds:
                shard, seq1 = shard.with_bounds(nb), None
        seq = seq1
        del seq1, cal
    if seq is None:
        seq = make_range(shard.start, shard.stop, B_total, C)
    Bl = shard.B
    V, F = len(seq["verts"]), len(seq["faces"])

    # ---- device-resident throughput: K fused iterations, inputs already in HBM
    model = model_from(seq)
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
    ms_rank = ev0.elapsed_time(ev1)
    ms = torch.tensor([ms_rank], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    value = B_total * args.steps / (ms / 1000.0)
    hist = fused.history()
    own = torch.tensor([float(np.mean(fused.compute_ms(args.warmup, args.warmup + args.steps)))], dtype=torch.float64,
                       device="cuda")
    rank_ms = [float(v) for v in allgather_equal(own, shard).reshape(-1).cpu().numpy()]

    # ---- per-kernel times (CUDA events on the launch stream) -> roofline of the dominant kernel
    # (the eager profile runs whole iterations: in a sharded run every rank does it, in lockstep like the timed loop)
    prof = fused.profile(5)
    fused.check_status()
    halo_used = fused.halo_mode
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
                "whole_step": {"algorithmic_bytes": ab["total"] * B_total,
                               "achieved": ab["total"] * B_total / (ms / args.steps / 1000.0) / 1e9,
                               "frac": ab["total"] * B_total / (ms / args.steps / 1000.0) / 1e9 / peak / world},
                "alu": alu_roofline(top, Bl, prof[top])}

    # ---- sharded == single GPU, bit for bit: 3 iterations from the initial poses on all ranks, then the whole
    #      sequence alone on rank 0 (which also gives the one-GPU time of this very configuration)
    extra = {}
    if world > 1 and not args.no_shard_check:
        from dynhor_b200.sharding import allgather_frames
        n_chk = 3
        m2 = model_from(seq)
        with FusedJointOpt(m2, lw, LR, n_chk, shard=shard, halo=args.halo) as f2:
            f2.run(n_chk)
            f2.check_status()
        pose = torch.cat([m2.rotations_object.detach().reshape(Bl, 6), m2.translations_object.detach().reshape(Bl, 3)], 1)
        pose = allgather_frames(pose, shard)
        del m2, f2, model, fused
        torch.cuda.empty_cache()
        if rank == 0:
            full = make_range(0, B_total, B_total, C)
            m1 = model_from(full)
            with FusedJointOpt(m1, lw, LR, n_chk + args.warmup + args.steps + 8) as f1:
                f1.run(n_chk)
                pose1 = torch.cat([m1.rotations_object.detach().reshape(B_total, 6),
                                   m1.translations_object.detach().reshape(B_total, 3)], 1)
                extra["shard_equals_single"] = bool(torch.equal(pose, pose1))
                extra["shard_check"] = {"iterations": n_chk, "frames": B_total,
                                        "max_abs_diff": float((pose - pose1).abs().max())}
                n1 = max(3, min(args.steps, 10))
                f1.run(3)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                f1.run(n1)
                e1.record()
                torch.cuda.synchronize()
                ms1 = e0.elapsed_time(e1) / n1
                extra["single_gpu_same_config"] = {"value": B_total / (ms1 / 1000.0), "unit": UNIT, "ms_per_step": ms1,
                                                   "steps": n1, "note": "the whole sequence on rank 0 alone"}
            del m1, f1, full
            torch.cuda.empty_cache()
        barrier()
    elif world == 1:
        del model, fused

    # ---- end to end through the public call with HOST buffers (H2D of inputs, D2H of results inside the timing).
    #      Every rank passes the whole-sequence list like run.py would; it holds real (pinned) tensors for the frames it
    #      can end up owning (its equal-count range, its cost-balanced range and a margin).
    if world > 1:
        eq = FrameShard(rank, world, B_total)
        lo, hi = max(0, min(eq.start, shard.start) - 32), min(B_total, max(eq.stop, shard.stop) + 32)
        hseq = seq if (lo, hi) == (shard.start, shard.stop) else make_range(lo, hi, B_total, C)
    else:
        lo, hseq = 0, seq
    params, nbytes = host_parameters(hseq, lo, B_total, C)
    del hseq
    faces_b = np.stack([seq["faces"]] * B_total)     # run.py:158
    e2e_iters = args.steps
    kw = dict(objvertices=seq["verts"], objfaces=faces_b, loss_weights=lw, lr=LR, board=None, halo=args.halo,
              balance="auto" if args.balance == "probe" else args.balance)   # the call's own policy: probe from 64 iterations on
    joint_optimize(params, num_iterations=2, **kw)
    # the call is one shot of ~45 ms of which several ms are host work (uploads, launches): timed args.e2e_reps times
    # from the same host buffers, the MEDIAN is reported (every repetition is listed in seconds_all)
    dts = []
    for _ in range(max(1, args.e2e_reps)):
        barrier()
        t0 = time.perf_counter()
        model2, evo = joint_optimize(params, num_iterations=e2e_iters, **kw)
        rot_h = model2.rotations_object.detach().cpu()
        tr_h = model2.translations_object.detach().cpu()
        torch.cuda.synchronize()
        dts.append(time.perf_counter() - t0)
    own = model2.frame_shard
    h2d = nbytes(params[own.start:own.stop]) + seq["verts"].nbytes + seq["faces"].astype(np.int32).nbytes
    d2h = rot_h.numel() * 4 + tr_h.numel() * 4 + 4 * 8 * e2e_iters
    dt_t = torch.tensor(dts, device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)      # per repetition: the slowest rank
    dts = [float(x) for x in dt_t.cpu()]
    dt = float(np.median(dts))
    e2e = {"value": B_total * e2e_iters / dt, "unit": UNIT, "h2d_bytes_per_step": h2d / e2e_iters,
           "d2h_bytes_per_step": d2h / e2e_iters, "iterations": e2e_iters, "seconds": dt,
           "seconds_all": [round(x, 5) for x in dts], "final_loss": evo["loss"][-1],
           "partition": "single" if world == 1 else ("count" if list(own.bounds) == list(FrameShard(rank, world, B_total).bounds)
                                                      else "probe")}
    if getattr(model2, "timing", None):     # DH_TIMING=1: synchronised phase times of this rank's call (diagnostics)
        e2e["phases_ms"] = {k: round(v, 2) for k, v in model2.timing}

    secondary = {}
    cb = None
    if rank == 0 and world == 1:
        if not args.no_secondary and not args.emulate_shard:
            del model2
            torch.cuda.empty_cache()
            try:
                secondary["dino"] = dino_numbers(max(3, min(args.steps, 20)))[0]
            except Exception as exc:  # the headline line must not die with a secondary workload
                secondary["dino"] = {"error": repr(exc)}
        if not args.no_cpu_baseline:
            cb, _ = cpu_baseline(C=C)
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if scaling == "strong" else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{workload}: {B_total} frames {H}x{W}"
                                   + (f", frame-sharded over {world} GPUs" if world > 1 else "")
                                   + f", {MESH} mesh (V={V}, F={F}), 256x256 ROI rendered 512x512 + 2x2 pool, {C} "
                                   f"correspondences per frame, lw_sil 1 / lw_smooth 10"
                                   + (f" / lw_corr {LW_CORR}" if C > 0 else "") + ", lr 1e-4",
                       "frames_total": B_total, "frames_this_rank": Bl, "correspondences_per_frame": C,
                       "parallelism": f"frame-shard x{world}",
                       "partition": {"by": args.balance if world > 1 else "single", "bounds": shard.bounds,
                                     "probe_ms_per_equal_range": probe_ms,
                                     "ms_per_probed_range_after_16_iterations": calib_ms},
                       "own_ms_per_step_by_rank": rank_ms,
                       "trajectory_period_frames": B_total,
                       "halo": halo_used,
                       "l2": "per-step working set (face-index maps + bins, > 1 MB/frame) exceeds the 126 MB L2; "
                             "no explicit flush", "cuda_graph": True},
            "roofline": roofline, "e2e": e2e, "clocks": clocks,
            "gpu_launches": (9 + (1 if C > 0 else 0)) * args.steps,
            "loss_first_last": [hist["loss"][0], hist["loss"][-1]],
            "iou_first_last": [hist["iou_object"][0], hist["iou_object"][-1]],
        }
        out.update(extra)
        if secondary:
            out["secondary"] = secondary
        if cb is not None:
            out["cpu_baseline"] = cb
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def alu_roofline(kernel, frames, kernel_ms):
    """Instruction roofline beside the HBM one: thread-instructions per frame of the kernel from the committed ncu
    capture of this round (profiles/alu.json, written by tools/ncu_summary.py) against the SM's issue rate."""
    path = os.path.join(ROOT, "profiles", "alu.json")
    try:
        a = json.load(open(path))[kernel]
    except Exception:
        return None
    sm_mhz = 1965.0
    try:
        sm_mhz = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["sm_max_mhz"])
    except Exception:
        pass
    peak = 148 * 4 * 32 * sm_mhz * 1e6            # thread-instructions / s: 148 SMs x 4 schedulers x 32 lanes
    ach = a["thread_inst_per_frame"] * frames / (kernel_ms / 1000.0)
    return {"thread_inst_per_frame": a["thread_inst_per_frame"], "achieved_inst_per_s": ach, "peak": peak,
            "frac": ach / peak, "issue_pct": a.get("issue_pct"), "lanes": a.get("lanes"),
            "source": a.get("source")}


def dino_numbers(steps, warmup=3, N=1000, D=384):
    """BASELINE configs[3]: DINO ViT-S/14 patch-feature matching, 1k templates x 300 frames, bf16 similarity GEMM
    (K = 1369*384) with the top-10 selection fused into it.  Returns the numbers of one JSON object."""
    import ctypes
    from dynhor_b200 import _lib, synth
    from dynhor_b200.dino_match import build_bank, dino_cos_topk
    Fm, P, k = 300, 1369, 10
    d = synth.make_dino_features(N, Fm, P, D, seed=0, device="cuda")
    tb = build_bank(d["templ"])
    fb = build_bank(d["frames"], d["masks"])
    if N > 2000:                      # the reference-sized bank: drop the 25 GB of fp32 features once the banks exist
        d["templ"] = d["templ"][:500].clone()
        torch.cuda.empty_cache()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(max(warmup, 3)):
        dino_cos_topk(fb, tb, k)
    times = []
    for _ in range(steps):
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
    plan = (ctypes.c_int32 * 8)()
    _lib.load().dh_dino_plan_info(N, Fm, P * D, plan)
    return {
        "metric": "DINO template-frame pairs scored / second (bf16 GEMM + fused top-k)", "value": N * Fm / (ms / 1e3),
        "unit": "pairs/s", "n_gpus": 1, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms,
        "higher_is_better": True, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": ("DINO ViT-S/14 patch-feature pose initialisation: 1000 templates x 300 frames, "
                                "P=1369, D=384, top-10 (BASELINE configs[3])") if (N, D) == (1000, 384) else
                               (f"DINO patch-feature pose initialisation at the reference's own size (dino.py:5 ViT-B/14, "
                                f"run.py:131-133): {N} templates x {Fm} frames, P={P}, D={D}, top-{k}"),
                   "l2": "256 MB flush between steps"},
        "roofline": {"bound": "hbm", "achieved": bytes_ / (ms / 1e3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": bytes_ / (ms / 1e3) / 1e9 / hbm, "peak_source": f"{kind} hbm_gbs",
                     "tensor": {"achieved_tflops": flops / (ms / 1e3) / 1e12, "peak_tflops": tf_peak,
                                "frac": flops / (ms / 1e3) / 1e12 / tf_peak}},
        "plan": dict(zip(["m_tiles", "n_pairs", "k_slices", "kblocks", "kb_per_slice", "cluster", "ctas_sized_for",
                          "ldc"], list(plan))),
        "planted_match_rank0": ok, "gpu_launches": 3 * steps}, d


def run_dino(args):
    """Secondary workload as its own line (`--workload dino`); `value` = template-frame pairs / second."""
    torch.cuda.set_device(0)
    out, d = dino_numbers(args.steps, args.warmup, args.dino_templates, args.dino_dim)
    # CPU baseline: the reference expression verbatim on a bounded sample of frames (and of templates at the big size)
    from oracle import dino_oracle
    nf, k = 2, 10
    N = d["templ"].shape[0]
    t0 = time.perf_counter()
    dino_oracle.dino_cos_topk(d["frames"][:nf].cpu(), d["masks"][:nf].cpu(), d["templ"].cpu(), k)
    cpu_s = (time.perf_counter() - t0) / nf
    out["cpu_baseline"] = {"value": N / cpu_s, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "reference",
                           "sample": f"{nf} frames x {N} templates, pose_initializtion.py:295-296 verbatim + topk"}
    print(json.dumps(out))


def run_stage1(args):
    """Secondary workload: the silhouette term of the per-frame pose initialisation,
    pose_initializtion.py:143-155,346-360 -- 1 - IoU + off-screen penalty on an anti_aliasing=False 256x256 render, one
    Adam group, lr 0.01 (configs/custom_shoes.yaml:12-13) -- for one candidate per frame of a custom_shoes-shaped
    sequence, all frames in one fused batch.  A step = one optimisation iteration of every candidate; `value` =
    candidate-iterations / second."""
    from dynhor_b200 import synth
    from dynhor_b200.jointopt import FusedJointOpt
    from dynhor_b200.pose_init import OFFSCREEN_WEIGHT, _Candidates, coarse_optimize
    B, lr = args.frames_per_gpu, 1e-2
    torch.cuda.set_device(0)
    seq = synth.make_sequence(B, H, W, mesh=MESH, seed=0, render_fn=gpu_render_fn, period=B)
    V, F = len(seq["verts"]), len(seq["faces"])
    lw = {"lw_sil_obj": 1.0, "lw_offscreen": OFFSCREEN_WEIGHT}
    unit = "candidate-iters/s"

    def candidates():
        return _Candidates(torch.from_numpy(seq["rot6d_init"]), torch.from_numpy(seq["T_init"]),
                           torch.from_numpy(seq["verts"]), torch.from_numpy(seq["faces"]),
                           torch.from_numpy(seq["K_roi"]), torch.from_numpy(seq["target_masks"]))

    model = candidates()
    fused = FusedJointOpt(model, lw, lr, args.steps + args.warmup + 16, stage1=True)
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.7)
    fused.run(args.warmup)
    torch.cuda.synchronize()
    sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    fused.run(args.steps)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    hist = fused.history()
    prof = fused.profile(5)
    fused.release()
    ab = {"raster": 12 * V + 6 * S * S + 8 * S * S, "backward": 5 * S * S + 8 * S * S + 24 * V,
          "project": 12 * V, "pose_update": 12 * V + 288, "total": 60 * V + 27 * S * S + 288}
    top = max(("project", "raster", "backward", "pose_update"), key=lambda k: prof[k])
    peak, peak_kind = measured_peaks()
    # end to end: host arrays in, poses + per-candidate losses out
    host = [torch.from_numpy(seq[k]).pin_memory() for k in ("target_masks", "rot6d_init", "T_init", "K_roi")]
    coarse_optimize(host[0], seq["verts"], seq["faces"], host[1], host[2], host[3], 2, lr)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = coarse_optimize(host[0], seq["verts"], seq["faces"], host[1], host[2], host[3], args.steps, lr)
    res = [out[k].cpu() for k in ("rotations", "translations", "losses")]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    h2d = sum(t.numel() * t.element_size() for t in host) + seq["verts"].nbytes + seq["faces"].nbytes
    d2h = sum(t.numel() * t.element_size() for t in res) + 4 * 8 * args.steps
    cb = None
    if not args.no_cpu_baseline:
        from oracle import stage1_oracle
        nf = max(2, min(os.cpu_count() or 1, 16))
        orc = stage1_oracle.Stage1Oracle(seq["target_masks"][:nf], seq["verts"], seq["faces"], seq["rot6d_init"][:nf],
                                         seq["T_init"][:nf], seq["K_roi"][:nf], lr=lr)
        t0 = time.perf_counter()
        orc.step()
        cpu_s = time.perf_counter() - t0
        cb = {"value": nf / cpu_s, "unit": unit, "cores": os.cpu_count(), "kind": "port",
              "sample": f"{nf} candidates x 1 iteration, oracle/stage1_oracle.py (pinned to runs of the reference's "
                        f"ObjTracker.coarse_forward), {cpu_s:.1f} s"}
    print(json.dumps({
        "metric": "stage-1 silhouette pose-initialisation candidate-iterations / second (fwd+bwd+Adam)",
        "value": B * args.steps / (ms / 1e3), "unit": unit, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"per-frame pose initialisation, silhouette term: {B} candidates (one per frame of a "
                               f"custom_shoes-shaped {H}x{W} sequence), {MESH} mesh (V={V}, F={F}), 256x256 render "
                               "without anti-aliasing, 1 - IoU + 100000 x off-screen, Adam lr 0.01",
                   "l2": "per-step working set exceeds the L2; no explicit flush", "cuda_graph": True},
        "roofline": {"bound": "hbm", "kernel": "k_" + top, "achieved": ab[top] * B / (prof[top] / 1e3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": ab[top] * B / (prof[top] / 1e3) / 1e9 / peak,
                     "peak_source": f"{peak_kind} hbm_gbs", "traffic": None,
                     "algorithmic_bytes_per_launch": ab[top] * B, "kernel_ms": prof[top], "kernel_ms_all": prof},
        "e2e": {"value": B * args.steps / dt, "unit": unit, "h2d_bytes_per_step": h2d / args.steps,
                "d2h_bytes_per_step": d2h / args.steps, "seconds": dt},
        "cpu_baseline": cb, "clocks": clocks, "gpu_launches": 9 * args.steps,
        "loss_first_last": [hist["loss"][0], hist["loss"][-1]],
        "iou_first_last": [hist["iou_object"][0], hist["iou_object"][-1]]}))


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
    ap.add_argument("--corr", type=int, default=None,
                    help="correspondences per frame (builder-defined reprojection term; 0 = the reference's two terms; "
                         "default 10000, 50000 in the strong-scaling sweep)")
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"],
                    help="auto: one GPU = BASELINE configs[1] (300 frames); several GPUs = strong scaling on configs[4] "
                         "(4096 frames, 50k correspondences, fixed total).  weak: --frames-per-gpu frames on every GPU")
    ap.add_argument("--frames-total", type=int, default=0, help="strong scaling: frames of the whole sequence (4096)")
    ap.add_argument("--balance", default="probe", choices=["probe", "count"],
                    help="frame ranges of equal measured cost (one probe pass) or of equal frame count")
    ap.add_argument("--no-shard-check", action="store_true",
                    help="skip the sharded == single-GPU comparison (and the one-GPU time of the same configuration)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the DINO numbers in the one-GPU line")
    ap.add_argument("--dino-templates", type=int, default=1000, help="--workload dino: templates (reference: 6000)")
    ap.add_argument("--dino-dim", type=int, default=384, help="--workload dino: feature channels (reference ViT-B: 768)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-reps", type=int, default=3, help="repetitions of the end-to-end call (median reported)")
    ap.add_argument("--mesh", default=MESH, choices=["uv50x100", "uv100x200"],
                    help="uv100x200 = the 20k-vertex mesh of BASELINE configs[2]")
    ap.add_argument("--camera", default=f"{H}x{W}", help="full-frame camera HxW (configs[2]: 1080x1920)")
    ap.add_argument("--emulate-shard", default="", help="r/w: time the frames of rank r of a w-GPU run on one GPU")
    ap.add_argument("--workload", default="jointopt", choices=["jointopt", "dino", "preprocess", "stage1"])
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"])
    args = ap.parse_args()
    MESH = args.mesh
    H, W = (int(v) for v in args.camera.split("x"))
    if args.workload == "dino":
        run_dino(args)
    elif args.workload == "preprocess":
        run_preprocess(args)
    elif args.workload == "stage1":
        run_stage1(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        run_ours(args)


if __name__ == "__main__":
    main()
