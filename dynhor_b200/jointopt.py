"""Drop-in for ObjTracker/jointopt.py on the B200 kernels.

Same public interface:
    Joint_Optimizer(translations_object, rotations_object, verts_object_og, faces_object, camintr_rois_object,
                    target_masks_object, int_scale_init=1.0, optimize_object_scale=False)     jointopt.py:15-62
        .get_verts_object()                                                                   jointopt.py:64-72
        .forward(loss_weights) -> (loss_dict, metric_dict)                                    jointopt.py:74-91
    joint_optimize(object_parameters, objvertices, objfaces, loss_weights, num_iterations, lr, board,
                   optimize_object_scale) -> (model, loss_evolution)                          jointopt.py:93-161
`Joint_Optimizer.forward` is the composable autograd path (CUDA renderer + torch ops for the scalar glue);
`joint_optimize` runs the fused kernels (forward + backward + Adam per iteration, CUDA-graph replayed, no
per-step host synchronisation) and fills `loss_evolution` / `board` after the loop from a device-side history.
"""
import ctypes
import os
import threading
from collections import defaultdict

import numpy as np

import torch
from torch import nn

from . import _lib
from .camera import compute_transformation_persp, tensorify
from .corr import CorrespondenceTerm, plan as corr_plan
from .geometry import matrix_to_rot6d, rot6d_to_matrix
from .losses import Losses
from .renderer import SilhouetteState, shared_faces
from .sharding import (FrameShard, PeerMailboxes, allgather_equal, allgather_frames, allreduce_sum_, balanced_bounds,
                       detect_shard, exchange_halo, frame_costs_from_blocks, rescale_costs, wants_probe)


class Joint_Optimizer(nn.Module):
    def __init__(self, translations_object, rotations_object, verts_object_og, faces_object, camintr_rois_object,
                 target_masks_object, int_scale_init=1.0, optimize_object_scale=False, correspondences=None,
                 corr_delta=1.0):
        """`correspondences` ([B,C,6], optional) and `corr_delta` belong to the builder-defined reprojection term
        (dynhor_b200/corr.py); without them the module is the reference's (jointopt.py:16-62)."""
        super().__init__()
        translation_init = translations_object.detach().clone()
        self.translations_object = nn.Parameter(translation_init, requires_grad=True)
        rotations_object = rotations_object.detach().clone()
        if rotations_object.shape[-1] == 3:
            rotations_object6d = matrix_to_rot6d(rotations_object)
        else:
            rotations_object6d = rotations_object
        self.rotations_object = nn.Parameter(rotations_object6d.detach().clone().contiguous(), requires_grad=True)
        self.register_buffer("verts_object_og", verts_object_og)
        init_scales = int_scale_init * torch.ones(1).float()
        self.optimize_object_scale = optimize_object_scale
        if optimize_object_scale:
            self.int_scales_object = nn.Parameter(init_scales, requires_grad=True)
        else:
            self.register_buffer("int_scales_object", init_scales)
        self.register_buffer("int_scale_object_mean", torch.ones(1).float())
        self.register_buffer("ref_mask_object", (target_masks_object > 0).float())
        self.register_buffer("keep_mask_object", (target_masks_object >= 0).float())
        self.register_buffer("camintr_rois_object", camintr_rois_object)
        self.register_buffer("faces_object", faces_object)
        if not torch.cuda.is_available():
            raise _lib.DynhorError("Joint_Optimizer needs a CUDA device (jointopt.py:56 calls .cuda(); dynhor_b200 "
                                   "has no CPU fallback)")
        self.cuda()
        self.losses = Losses(ref_mask_object=self.ref_mask_object, keep_mask_object=self.keep_mask_object,
                             camintr_rois_object=self.camintr_rois_object)
        self.corr_delta = float(corr_delta)
        self.corr_term = None
        if isinstance(correspondences, CorrespondenceTerm):     # built by the caller (joint_optimize: on its upload stream)
            self.corr_term = correspondences
        elif correspondences is not None:
            self.corr_term = CorrespondenceTerm(correspondences.cuda(), self.camintr_rois_object,
                                                image_size=int(target_masks_object.shape[-1]), delta=corr_delta)

    def get_verts_object(self):
        rotations_object = rot6d_to_matrix(self.rotations_object)
        return compute_transformation_persp(meshes=self.verts_object_og, translations=self.translations_object,
                                            rotations=rotations_object,
                                            intrinsic_scales=self.int_scales_object.abs())

    def forward(self, loss_weights=None):
        """If a loss weight is zero, that loss isn't computed."""
        loss_dict = {}
        metric_dict = {}
        verts_object = self.get_verts_object()
        if loss_weights is None or (loss_weights["lw_smooth_obj"] > 0):
            loss_dict.update(self.losses.compute_smooth_loss(verts_object))
        if loss_weights is None or loss_weights["lw_sil_obj"] > 0:
            sil_loss_dict, sil_metric_dict = self.losses.compute_sil_loss(verts=verts_object,
                                                                          faces=self.faces_object)
            loss_dict.update(sil_loss_dict)
            metric_dict.update(sil_metric_dict)
        if self.corr_term is not None and (loss_weights is None or loss_weights.get("lw_corr_obj", 0) > 0):
            loss_dict["loss_corr_obj"] = self.corr_term.loss(rot6d_to_matrix(self.rotations_object),
                                                             self.translations_object,
                                                             self.int_scales_object.abs())
        return loss_dict, metric_dict


class FusedJointOpt:
    """The fused iteration (dh_jointopt_run) bound to a Joint_Optimizer's parameters, updated in place."""

    def __init__(self, model, loss_weights, lr, max_iters, shard=None, group=None, nchunks=None, keep_sum=None,
                 exchange=True, halo="p2p", corr_on=None, halo_timeout_ms=0, stage1=False, start_step=0):
        """halo: how boundary poses (and, with optimize_object_scale, the partial scale gradients) travel between
        ranks when the sequence is sharded.  "p2p" (default): CUDA-IPC mailboxes written by the kernels themselves
        over NVLink, no host work per iteration; falls back to "nccl" when the ranks cannot map each other's memory.
        "nccl": one grouped send/recv (+ one all_gather for the scale) per iteration driven from the host (also the
        path the gloo CPU tests exercise).  exchange=False: the caller fills self.halo by hand and passes keep_sum
        (single-process shard emulation; optimize_object_scale then needs all emulated shards stepped through
        `step_emulated`).  corr_on: whether the correspondence term is active -- must be the same on every rank
        (default: this model has records and lw_corr_obj > 0).
        stage1=True: the silhouette term of the per-frame pose initialisation instead of jointopt's losses
        (DH_LOSS_STAGE1, include/dynhor_b200.h): loss_weights = {"lw_sil_obj": weight of 1 - IoU,
        "lw_offscreen": weight of the off-screen penalty}, no anti-aliasing, one Adam group, frames independent.
        start_step: iterations already done on these parameters by an earlier plan (a re-partition in the middle of
        a run): the iteration counter -- Adam's t, the history row, the mailbox ticks -- continues from there."""
        lib = _lib.load()
        self.model, self.group = model, group
        rot, tr = model.rotations_object, model.translations_object
        if not rot.is_cuda:
            raise _lib.DynhorError("FusedJointOpt needs CUDA parameters (no CPU fallback)")
        dev = rot.device
        B = rot.shape[0]
        self.shard = shard if shard is not None else FrameShard(0, 1, B)
        if self.shard.B != B:
            raise ValueError(f"model holds {B} frames but the shard owns {self.shard.B}")
        assert rot.is_contiguous() and tr.is_contiguous() and rot.dtype == torch.float32
        masks = model.ref_mask_object + model.keep_mask_object - 1.0  # back to {-1, 0, 1} (jointopt.py:50-53)
        S = int(masks.shape[-1])
        assert masks.shape == (B, S, S), "target masks must be [B,S,S]"
        verts = model.verts_object_og.detach().contiguous().float()
        assert verts.ndim == 2 and verts.shape[-1] == 3, "Invalid shape for vertices"
        V = verts.shape[0]
        faces = shared_faces(model.faces_object)
        self.stage1 = bool(stage1)
        self.sil = SilhouetteState(B, V, faces, model.camintr_rois_object, S, not self.stage1, orig_size=1.0)
        self.verts = verts
        st = _lib.stream_ptr()
        # static inputs
        self.mask_tri = torch.empty(B, S, S, dtype=torch.int8, device=dev)
        keep = torch.zeros(1, dtype=torch.int64, device=dev)
        _lib.check(lib.dh_masks_prepare(_lib.ptr(masks.contiguous().float()), _lib.ptr(self.mask_tri), _lib.ptr(keep),
                                        masks.numel(), st), "dh_masks_prepare")
        del masks
        self.exchange = exchange  # False: the caller fills self.halo by hand (single-process shard emulation)
        # builder-defined correspondence term: on only with records AND a positive weight, on EVERY rank alike
        lw_corr = float(loss_weights.get("lw_corr_obj", 0.0))
        has_corr = getattr(model, "corr_term", None) is not None and lw_corr > 0
        self.corr_on = has_corr if corr_on is None else bool(corr_on)
        if self.corr_on and not has_corr:
            raise ValueError("corr_on=True but this rank's model has no correspondences / lw_corr_obj")
        # one fused all-reduce for the sequence-wide constants: sum(keep), sum(w), and the ranks' view of corr_on
        # (a rank that disagreed would otherwise hang the others in a mismatched collective later)
        # Single rank: the records may still be on their way (joint_optimize uploads them on a side stream).  Then the
        # kernels read the sum of weights from device memory (dh_corr.w_sum_dev) and the first iteration runs as two
        # halves with the wait for the upload between them (_run): its silhouette kernels overlap the copy.
        # Sharded: the same, with the all-reduce of the sum of weights queued behind the upload on the upload stream.
        self._corr_event, self._w_sum_dev = None, None
        if self.corr_on and model.corr_term.ready is not None:
            if self.shard.world == 1:
                self._corr_event, model.corr_term.ready = model.corr_term.ready, None
                self._w_sum_dev = model.corr_term.w_local
            elif exchange:
                self._w_sum_dev = torch.zeros(1, dtype=torch.float64, device=dev)
                self._corr_event = "pending"     # the all-reduce is queued at the end of __init__, behind the halo exchange
            else:
                model.corr_term.wait_ready()
        consts = torch.zeros(4, dtype=torch.float64, device=dev)
        consts[0] = keep[0].to(torch.float64)
        if self.corr_on and self._corr_event is None:
            consts[1] = model.corr_term.w_local.reshape(()).to(torch.float64)
        consts[2] = 1.0 if self.corr_on else 0.0
        consts[3] = 1.0
        if exchange and self.shard.world > 1:
            allreduce_sum_(consts, self.shard, group)
        # the host needs them (keep_sum enters the kernels by value): copied back asynchronously, read after the
        # allocations below -- the host works on while the masks are still crossing PCIe
        consts_pin, consts_ev = None, None
        if keep_sum is None or (self.corr_on and self._corr_event is None) or exchange:
            consts_pin = torch.empty(4, dtype=torch.float64, pin_memory=True)
            consts_pin.copy_(consts, non_blocking=True)
            consts_ev = torch.cuda.Event()
            consts_ev.record()
        self.keep_local = keep
        self.moments = torch.empty(12, dtype=torch.float64, device=dev)
        _lib.check(lib.dh_mesh_moments(_lib.ptr(verts), V, _lib.ptr(self.moments), st), "dh_mesh_moments")
        # optimiser state / history
        self.scale = model.int_scales_object
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)  # noqa: E731
        self.m_rot, self.v_rot, self.m_tr, self.v_tr, self.mv_scale = z(B, 6), z(B, 6), z(B, 3), z(B, 3), z(2)
        self.step = z(1, dt=torch.int32)
        self.start_step = int(start_step)
        if self.start_step:
            self.step.fill_(self.start_step)
        self.status = z(1, dt=torch.int32)
        self.scale_part = z(2, dt=torch.int64)
        self.max_iters = int(max_iters)
        self.hist = z(self.max_iters + 1, 4, dt=torch.float64)   # row max_iters: evaluate() / grads()
        self.iter_ns = z(max(self.max_iters, 1), 2, dt=torch.int64)
        self.halo = z(2, 9)
        self.edge = z(2, 9)
        nchunks = nchunks or int(os.environ.get("DH_BWD_CHUNKS", "0"))  # tuning knob
        self.nchunks = int(nchunks) if nchunks else lib.dh_jointopt_default_chunks(B, self.sil.F)
        sizes = (ctypes.c_int64 * 5)()
        _lib.check(lib.dh_jointopt_scratch_bytes(B, self.nchunks, sizes), "dh_jointopt_scratch_bytes")
        self.scratch = [torch.zeros(int(n), dtype=torch.uint8, device=dev) for n in sizes]
        self.loss_weights = dict(loss_weights)
        consts_h = None
        if consts_ev is not None:
            consts_ev.synchronize()
            consts_h = consts_pin.numpy().copy()
        if exchange and self.shard.world > 1 and consts_h[2] not in (0.0, consts_h[3]):
            raise _lib.DynhorError("the correspondence term is active on some ranks only: pass the same "
                                   "loss_weights / correspondences to every rank")
        self.keep_sum = float(keep_sum) if keep_sum is not None else float(consts_h[0])
        p = _lib.DhJointOpt()
        p.sil = self.sil.c
        p.verts_og, p.mask_tri = verts.data_ptr(), self.mask_tri.data_ptr()
        p.rot6d, p.trans, p.scale = rot.data_ptr(), tr.data_ptr(), self.scale.data_ptr()
        p.adam_m_rot, p.adam_v_rot = self.m_rot.data_ptr(), self.v_rot.data_ptr()
        p.adam_m_trans, p.adam_v_trans = self.m_tr.data_ptr(), self.v_tr.data_ptr()
        p.adam_mv_scale = self.mv_scale.data_ptr()
        p.step, p.hist, p.max_iters = self.step.data_ptr(), self.hist.data_ptr(), self.max_iters
        p.status, p.scale_part = self.status.data_ptr(), self.scale_part.data_ptr()
        p.iter_ns = self.iter_ns.data_ptr()
        p.halo_timeout_ms = int(halo_timeout_ms)
        p.halo_prev = self.halo[0].data_ptr() if self.shard.has_prev else None
        p.halo_next = self.halo[1].data_ptr() if self.shard.has_next else None
        p.B_total = self.shard.B_total
        p.rank, p.world = self.shard.rank, self.shard.world
        p.keep_sum = self.keep_sum
        p.lw_sil = float(loss_weights.get("lw_sil_obj", 0.0))
        p.lw_smooth = float(loss_weights.get("lw_smooth_obj", 0.0))
        p.lr = float(lr)
        p.optimize_scale = int(bool(getattr(model, "optimize_object_scale", False)))
        p.moments = self.moments.data_ptr()
        (p.Rmat, p.smooth_terms, p.loss_counts, p.partials, p.frame_terms) = [b.data_ptr() for b in self.scratch]
        p.nchunks = self.nchunks
        if self.stage1:
            if p.lw_smooth > 0 or p.optimize_scale or self.corr_on or self.shard.world > 1:
                raise ValueError("stage-1 mode: silhouette IoU + off-screen penalty only, frames are independent")
            self.offscreen, self.frame_coef = z(B, 16), z(B, 2)
            p.loss_mode, p.lw_offscreen = _lib.LOSS_STAGE1, float(loss_weights.get("lw_offscreen", 0.0))
            p.offscreen, p.frame_coef = self.offscreen.data_ptr(), self.frame_coef.data_ptr()
        if self.corr_on:
            ct = model.corr_term
            cp = corr_plan(B, ct.records.shape[1])
            self.corr_partials = z(B, cp["nslots"], 16)
            p.corr.records, p.corr.C, p.corr.nslots = ct.records.data_ptr(), ct.records.shape[1], cp["nslots"]
            if self._corr_event is not None:
                self.corr_w_sum = None          # known on the device only
                self._w_sum_dev.record_stream(torch.cuda.current_stream())
                p.corr.w_sum_dev, p.corr.w_sum = self._w_sum_dev.data_ptr(), 0.0
            else:
                self.corr_w_sum = float(consts_h[1]) if (exchange and consts_h is not None) else ct.w_sum
                p.corr.w_sum = self.corr_w_sum
            p.corr.delta, p.corr.lw_corr = ct.delta, lw_corr
            p.corr.partials = self.corr_partials.data_ptr()
        self.p = p
        sharded = self.shard.world > 1
        self.halo_mode = halo if (sharded and exchange) else "none"
        self._mail, self._keep, self._handle = None, [], None
        self._sync_halo()
        if self.halo_mode == "p2p":
            self._setup_p2p()
        if self._corr_event == "pending":
            # sum of the correspondence weights over the ranks, on the upload stream behind the records' copy.  Queued
            # last: collectives run in issue order, and the halo exchange above must not wait for the upload.
            cur, up = torch.cuda.current_stream(), upload_stream(dev, deferred=True)
            up.wait_stream(cur)
            with torch.cuda.stream(up):
                self._w_sum_dev.copy_(model.corr_term.w_local.reshape(1))
                allreduce_sum_(self._w_sum_dev, self.shard, group)
                self._w_sum_dev.record_stream(up)
                self._corr_event = up.record_event()
            model.corr_term.ready = None
        # the shared scale's gradient under sharding: exact partial sums, added up in rank order on every rank
        if p.optimize_scale and sharded:
            p.scale_mode = _lib.SCALE_P2P if self.halo_mode == "p2p" else _lib.SCALE_DEFERRED
        else:
            p.scale_mode = _lib.SCALE_LOCAL

    # -- sharding ---------------------------------------------------------------------------------------------
    def _sync_halo(self):
        """Swap boundary-frame poses with the neighbouring ranks (9 floats each way)."""
        if self.shard.world == 1 or not self.exchange:
            return
        rot, tr = self.model.rotations_object.detach(), self.model.translations_object.detach()
        self.edge[0, :6] = rot[0].reshape(6)
        self.edge[0, 6:] = tr[0].reshape(3)
        self.edge[1, :6] = rot[-1].reshape(6)
        self.edge[1, 6:] = tr[-1].reshape(3)
        exchange_halo(self.edge[0], self.edge[1], self.shard, self.halo[0], self.halo[1], self.group)

    def _setup_p2p(self):
        """Bind this run to the process-wide mailboxes (sharding.PeerMailboxes: allocated and IPC-mapped once per
        process), reserve its tick range and seed this rank's own slots with the initial halo (already exchanged
        once through the process group).  No barrier: a neighbour's first publish goes to another slot."""
        mail = PeerMailboxes.get(self.shard, self.group)
        if mail is None:          # no peer access between the ranks' devices: host-driven exchange
            self.halo_mode = "nccl"
            return
        self._mail = mail
        base = mail.reserve(self.max_iters)
        for side in (0, 1):
            if (side == 0 and self.shard.has_prev) or (side == 1 and self.shard.has_next):
                self._keep.append(mail.seed(side, base + self.start_step, self.halo[side]))
        p = self.p
        p.mailbox, p.tick_base = mail.mailbox, base
        p.peer_prev = mail.peers[self.shard.rank - 1] if self.shard.has_prev else None
        p.peer_next = mail.peers[self.shard.rank + 1] if self.shard.has_next else None
        for r in range(self.shard.world):
            p.peers[r] = mail.peers[r]

    # -- execution --------------------------------------------------------------------------------------------
    def run(self, n_iters, use_graph=True):
        """n_iters fused iterations on the current stream; no host synchronisation (single rank / p2p).  Dispatched
        as the torch.library op dynhor::jointopt_run, which declares the parameters it updates in place."""
        from . import ops
        if self._handle is None:
            self._handle = ops.register_fused(self)
        torch.ops.dynhor.jointopt_run(self.model.rotations_object.detach(), self.model.translations_object.detach(),
                                      self.scale.detach(), self._handle, int(n_iters), bool(use_graph))

    def _wait_corr(self):
        if self._corr_event is not None:
            torch.cuda.current_stream().wait_event(self._corr_event)
            self._corr_event = None

    def _run(self, n_iters, use_graph=True):
        lib = _lib.load()
        n_iters = int(n_iters)
        if self._corr_event is not None and n_iters > 0:
            # first iteration: everything but the correspondence kernel, then wait for the records, then the rest
            _lib.check(lib.dh_jointopt_run_part(ctypes.byref(self.p), 1, _lib.stream_ptr()), "dh_jointopt_run_part")
            self._wait_corr()
            _lib.check(lib.dh_jointopt_run_part(ctypes.byref(self.p), 2, _lib.stream_ptr()), "dh_jointopt_run_part")
            n_iters -= 1
            if self.halo_mode == "nccl":
                if self.p.scale_mode == _lib.SCALE_DEFERRED:
                    self.apply_scale(allgather_equal(self.scale_part, self.shard, self.group))
                self._sync_halo()
        if self.halo_mode != "nccl":
            _lib.check(lib.dh_jointopt_run(ctypes.byref(self.p), int(n_iters), int(use_graph), _lib.stream_ptr()),
                       "dh_jointopt_run")
            return
        for _ in range(int(n_iters)):
            _lib.check(lib.dh_jointopt_run(ctypes.byref(self.p), 1, int(use_graph), _lib.stream_ptr()),
                       "dh_jointopt_run")
            if self.p.scale_mode == _lib.SCALE_DEFERRED:
                self.apply_scale(allgather_equal(self.scale_part, self.shard, self.group))
            self._sync_halo()

    def apply_scale(self, parts):
        """DH_SCALE_DEFERRED: Adam step of the shared scale from the exact partial gradients of all ranks
        ([world, 2] int64 on the device, rank order: every rank's `scale_part` after its iteration)."""
        parts = parts.contiguous()
        assert parts.shape == (self.shard.world, 2) and parts.dtype == torch.int64 and parts.is_cuda
        _lib.check(_lib.load().dh_scale_apply(ctypes.byref(self.p), _lib.ptr(parts), self.shard.world,
                                              _lib.stream_ptr()), "dh_scale_apply")

    def check_status(self):
        """Raise if a kernel gave up waiting for a neighbour (synchronises)."""
        code = int(self.status.item())
        if code == _lib.STATUS_HALO_TIMEOUT:
            raise _lib.DynhorError(f"{self.shard}: timed out waiting for a neighbouring rank's boundary pose / scale "
                                   "gradient (a rank died or fell behind by more than halo_timeout_ms)")
        if code:
            raise _lib.DynhorError(f"{self.shard}: device status {code}")

    def grads(self):
        """Gradients of the weighted loss for the current parameters (no update)."""
        B = self.shard.B
        dev = self.step.device
        self._wait_corr()
        g_rot = torch.empty(B, 3, 2, device=dev)
        g_tr = torch.empty(B, 1, 3, device=dev)
        g_s = torch.zeros(1, device=dev)
        _lib.check(_lib.load().dh_jointopt_grads(ctypes.byref(self.p), _lib.ptr(g_rot), _lib.ptr(g_tr), _lib.ptr(g_s),
                                                 _lib.stream_ptr()), "dh_jointopt_grads")
        return g_rot, g_tr, g_s

    def evaluate(self):
        """Losses of the current parameters -> dict of one-element lists (synchronises)."""
        self._wait_corr()
        _lib.check(_lib.load().dh_jointopt_eval(ctypes.byref(self.p), _lib.stream_ptr()), "dh_jointopt_eval")
        return self._rows_to_dict(self.hist[self.max_iters:self.max_iters + 1].clone())

    def probe(self, nblocks):
        """Milliseconds of the heavy kernels of one iteration for `nblocks` equal blocks of this rank's frames, plus
        the correspondence kernel over all of them as the last entry (dh_jointopt_probe; parameters untouched)."""
        self._wait_corr()
        ms = (ctypes.c_float * (int(nblocks) + 1))()
        _lib.check(_lib.load().dh_jointopt_probe(ctypes.byref(self.p), int(nblocks), ms, _lib.stream_ptr()),
                   "dh_jointopt_probe")
        return np.asarray(list(ms), np.float64)

    def _rows_to_dict(self, rows):
        rows = rows.clone()
        if self.exchange:
            allreduce_sum_(rows, self.shard, self.group)
        rows = rows.cpu().numpy()
        lw = self.loss_weights
        evo = defaultdict(list)
        if self.stage1:   # rows: sum of off-screen penalties, sum of (1 - IoU), mean IoU
            for r in rows:
                evo["offscreen"].append(float(r[0]))
                evo["iou_loss"].append(float(r[1]))
                evo["iou_object"].append(float(r[2]))
                evo["loss"].append(float(r[1]) * lw.get("lw_sil_obj", 0.0) + float(r[0]) * lw.get("lw_offscreen", 0.0))
            return dict(evo)
        for r in rows:
            total = 0.0
            if lw.get("lw_smooth_obj", 0) > 0:
                evo["loss_smooth_obj"].append(float(r[0]))
                total += float(r[0]) * lw["lw_smooth_obj"]
            if lw.get("lw_sil_obj", 0) > 0:
                evo["loss_sil_obj"].append(float(r[1]))
                total += float(r[1]) * lw["lw_sil_obj"]
                evo["iou_object"].append(float(r[2]))
            if self.corr_on:
                evo["loss_corr_obj"].append(float(r[3]))
                total += float(r[3]) * lw["lw_corr_obj"]
            evo["loss"].append(total)
        return dict(evo)

    def history(self):
        """loss_evolution of all iterations run so far (synchronises once)."""
        n = min(int(self.step.item()), self.max_iters)
        self.check_status()
        return self._rows_to_dict(self.hist[:n])

    def compute_ms(self, first=0, last=None):
        """This rank's own time per iteration in milliseconds, waits on the neighbours excluded (device clock between
        the end of the halo wait and the end of the iteration), for iterations [first, last) -> float64 array.
        Synchronises."""
        n = min(int(self.step.item()), self.max_iters)
        t = self.iter_ns[first:n if last is None else min(last, n)].cpu().numpy()
        return (t[:, 1] - t[:, 0]) * 1e-6

    def frame_losses(self):
        """Per-frame terms of the LAST evaluated iteration (the parameters before its update, like the `losses` vector
        the reference keeps at pose_initializtion.py:352-356) -> dict of [B] float64 tensors (synchronises nothing)."""
        ft = self.scratch[4].view(torch.float64).view(-1, 8)
        lw = self.loss_weights
        out = {"iou": ft[:, 1].clone(), "offscreen": ft[:, 5].clone()}
        if self.stage1:
            out["loss"] = lw.get("lw_sil_obj", 0.0) * (1.0 - ft[:, 1]) + lw.get("lw_offscreen", 0.0) * ft[:, 5]
        return out

    KERNELS = ("pose_prep", "project", "setup_bin", "raster", "backward", "pose_update", "finalize", "corr")

    def profile(self, n_iters=5):
        """Average per-kernel milliseconds over n_iters real iterations (CUDA events on the launch stream)."""
        self._wait_corr()
        ms = (ctypes.c_float * 8)()
        _lib.check(_lib.load().dh_jointopt_profile(ctypes.byref(self.p), int(n_iters), ms, _lib.stream_ptr()),
                   "dh_jointopt_profile")
        return {k: float(ms[i]) for i, k in enumerate(self.KERNELS)}

    def release(self):
        """Drop the cached graph executables of this plan.  The mailboxes stay (process-wide, reused)."""
        if getattr(self, "p", None) is not None:
            _lib.load().dh_jointopt_release(ctypes.byref(self.p))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.release()
        return False

    def __del__(self):
        try:
            self.release()
        except Exception:  # pragma: no cover  (interpreter shutdown)
            pass


def _stack_frames(frames, key, pick=None, dtype=None):
    """One tensor [n, ...] on the GPU from the per-frame entries `frames[i][key]` ([1, ...] each).  Device tensors are
    concatenated on the device; pinned host tensors are copied asynchronously straight into their rows of the
    result; pageable host tensors go through ONE pinned staging buffer and ONE host-to-device copy."""
    ts = [f[key] if pick is None else pick(f[key]) for f in frames]
    t0 = ts[0]
    if t0.is_cuda:
        out = torch.cat(ts)
    elif t0.is_pinned() and t0.numel() * t0.element_size() >= 65536:
        out = torch.empty((sum(int(t.shape[0]) for t in ts),) + tuple(t0.shape[1:]), dtype=t0.dtype, device="cuda")
        # (is_pinned() is a driver query of a few microseconds: asked of the first, the middle and the last frame, not
        # of all of them)
        if t0.shape[0] == 1 and all(t.shape == t0.shape and t.is_contiguous() and t.dtype == t0.dtype for t in ts) and \
                ts[len(ts) // 2].is_pinned() and ts[-1].is_pinned():
            _upload_rows(out, ts)      # one batched driver call for all the frames
        else:
            row = 0
            for t in ts:
                out[row:row + t.shape[0]].copy_(t, non_blocking=True)
                row += t.shape[0]
    else:
        n = sum(int(t.shape[0]) for t in ts)
        stage = torch.empty((n,) + tuple(t0.shape[1:]), dtype=t0.dtype, pin_memory=True)
        torch.cat(ts, out=stage)
        out = stage.cuda(non_blocking=True)
    return out if dtype is None or out.dtype == dtype else out.to(dtype)


_upload_stream = {}


def upload_stream(dev, deferred=False):
    """The per-device side streams of the batched uploads: one the compute stream waits for right away, and one for
    the deferred upload of the correspondence records (so that a later ordinary upload does not wait behind it)."""
    dev = torch.device(dev) if not isinstance(dev, int) else torch.device("cuda", dev)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    up = _upload_stream.get((idx, bool(deferred)))
    if up is None:
        up = _upload_stream[(idx, bool(deferred))] = torch.cuda.Stream(idx)
    return up


def _upload_rows(out, rows):
    """rows: equal-shaped contiguous pinned host tensors -> out[i] (dh_upload_rows: one cudaMemcpyBatchAsync instead of
    len(rows) copy calls).  Batched copies need a real stream, torch's default one is the legacy stream: they go through
    a per-device side stream the current stream then waits for (when the upload stream itself is current -- the
    deferred correspondence upload of joint_optimize -- nobody waits here)."""
    dev = out.device
    cur = torch.cuda.current_stream(dev)
    up = upload_stream(dev)
    if cur == up or cur == upload_stream(dev, deferred=True):
        up = cur
        n = len(rows)
        ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in rows])
        _lib.check(_lib.load().dh_upload_rows(ctypes.c_void_p(out.data_ptr()), ptrs,
                                              rows[0].numel() * rows[0].element_size(), n,
                                              ctypes.c_void_p(up.cuda_stream)), "dh_upload_rows")
        return
    up.wait_stream(cur)                       # `out` may reuse memory the current stream is still working on
    n = len(rows)
    ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in rows])
    _lib.check(_lib.load().dh_upload_rows(ctypes.c_void_p(out.data_ptr()), ptrs, rows[0].numel() * rows[0].element_size(),
                                          n, ctypes.c_void_p(up.cuda_stream)), "dh_upload_rows")
    out.record_stream(up)
    cur.wait_stream(up)


class _SharedFaces:
    """run.py:158 stacks one identical face list per frame ([B,F,3] int64: 72 MB at 300 frames, 1 GB at 4096).  The
    list is uploaded ONCE as int32 [F,3]; the rows are compared with row 0 on the host only for the frames a rank
    actually owns (so the ranks of a sharded run split the check), skipped altogether for a broadcast view."""

    def __init__(self, objfaces):
        f = objfaces if torch.is_tensor(objfaces) else torch.from_numpy(np.asarray(objfaces))
        self.rows = f if (f.ndim == 3 and f.shape[0] > 1 and f.stride(0) != 0) else None
        one = f[0] if f.ndim == 3 else f
        if one.ndim != 2 or one.shape[-1] != 3:
            raise AssertionError("Invalid shape for faces")
        self.first = one
        self.dev = one.to(torch.int32).contiguous().cuda()
        self.checked = []
        self.pending, self.bad = [], False

    def check(self, start, stop):
        """Starts the comparison of rows [start, stop) on a worker thread (72 MB of host reads at 300 frames: it runs
        beside the uploads and the kernel launches); `join` -- called before joint_optimize returns anything -- raises
        if a row differed."""
        if self.rows is None:
            return
        stop = min(stop, self.rows.shape[0])
        for a, b in self.checked:          # only what no earlier call covered
            if a <= start < b:
                start = b
            if a < stop <= b:
                stop = a
        if start >= stop:
            return
        self.checked.append((start, stop))

        def work():
            if not torch.equal(self.rows[start:stop], self.first.expand(stop - start, -1, -1)):
                self.bad = True

        t = threading.Thread(target=work, daemon=True)
        t.start()
        self.pending.append(t)

    def join(self):
        for t in self.pending:
            t.join()
        self.pending = []
        if self.bad:
            raise NotImplementedError("dynhor_b200 renders one mesh topology for all frames (run.py:158 stacks "
                                      "identical faces); per-frame face lists are not supported")


def joint_optimize(object_parameters, objvertices=None, objfaces=None, loss_weights=None, num_iterations=400,
                   lr=1e-4, board=None, optimize_object_scale=False, shard=None, use_graph=True, halo="p2p",
                   corr_delta=1.0, balance="auto", _rebalance_to=None):
    """jointopt.py:93-161.  Extra keywords: `shard` (a sharding.FrameShard): when given (or when torch.distributed
    is initialised with more than one rank) `object_parameters` is the full sequence and this rank optimises its
    contiguous frame range; the returned model holds the gathered poses of ALL frames on every rank.
    `balance` ("auto" | "probe" | "count"): how the sequence is cut when no shard is given -- by measured per-frame
    cost (one untimed + one timed pass of the heavy kernels over equal-count ranges, then ranges of equal cost) or by
    frame count.  The probe costs about as much as five iterations and buys about a tenth of every iteration (measured,
    8 GPUs, 4096 frames: 3.2 instead of 3.6 ms): "auto" probes from 64 iterations on.
    (_rebalance_to: bounds the mid-run re-partition must move to -- tests only.)"""
    if not torch.cuda.is_available():
        raise _lib.DynhorError("joint_optimize needs a CUDA device (dynhor_b200 has no CPU fallback)")
    if loss_weights is None:
        loss_weights = {"lw_sil_obj": 1.0, "lw_smooth_obj": 1.0}
    B_total = len(object_parameters)
    marks = []

    def mark(label):   # DH_TIMING=1: synchronised phase times in model.timing (diagnostics; off: no synchronisation)
        if os.environ.get("DH_TIMING"):
            import time
            torch.cuda.synchronize()
            marks.append((label, time.perf_counter()))

    mark("start")
    auto = shard is None
    shard = detect_shard(B_total) if shard is None else shard
    verts_object_og = tensorify(objvertices).cuda()
    faces = _SharedFaces(objfaces)
    faces_dev = faces.dev
    # the correspondence term is on for everybody or for nobody: decided from the FULL list, not this rank's slice
    corr_on = loss_weights.get("lw_corr_obj", 0) > 0 and all("correspondences" in obj for obj in object_parameters)

    def frames_on_device(sh, key, have=None, **kw):
        """[sh.B, ...] tensor of per-frame entry `key`; `have` = (tensor, shard) of an earlier upload whose frames are
        reused (device-side slice), only the frames it lacks are uploaded."""
        if have is None or have[0] is None:
            return _stack_frames(sh.slice(object_parameters), key, **kw)
        old, s0 = have
        lo, hi = max(sh.start, s0.start), min(sh.stop, s0.stop)
        if lo >= hi:
            return _stack_frames(sh.slice(object_parameters), key, **kw)
        parts = []
        if sh.start < lo:
            parts.append(_stack_frames(object_parameters[sh.start:lo], key, **kw))
        parts.append(old[lo - s0.start:hi - s0.start])
        if hi < sh.stop:
            parts.append(_stack_frames(object_parameters[hi:sh.stop], key, **kw))
        return torch.cat(parts) if len(parts) > 1 else parts[0]

    def start_corr_upload(sh):
        """Begin uploading the correspondence records of the frames of `sh` on the upload stream, without making the
        compute stream wait: (records tensor, shard) for build's `pre_corr`, or None when there is nothing to do."""
        local = sh.slice(object_parameters)
        if not corr_on or local[0]["correspondences"].is_cuda:
            return None
        cur, up = torch.cuda.current_stream(), upload_stream(torch.cuda.current_device(), deferred=True)
        up.wait_stream(cur)
        with torch.cuda.stream(up):
            rec = _stack_frames(local, "correspondences")
            rec.record_stream(cur)
        return rec, sh

    def build(sh, with_corr, reuse=None, pre_corr=None):
        """Model of the frames of `sh`.  Stage 1 hands over CUDA tensors (pose_initializtion.py:460-471); host tensors
        are accepted too.  reuse = (model, shard) of an earlier build: frames it already holds stay on the device.
        pre_corr = start_corr_upload's result for an earlier range: only the frames it lacks are uploaded."""
        local = sh.slice(object_parameters)
        faces.check(sh.start, sh.stop)
        trans = _stack_frames(local, "translations")
        rots = _stack_frames(local, "rotations")
        K = _stack_frames(local, "K_roi")[:, 0]
        m0, s0 = reuse if reuse is not None else (None, None)
        old_masks = None if m0 is None else m0.ref_mask_object + m0.keep_mask_object - 1.0
        masks = frames_on_device(sh, "target_masks", (old_masks, s0), dtype=torch.float32)
        corr = None
        if with_corr and corr_on:
            old_corr = m0.corr_term.records if (m0 is not None and getattr(m0, "corr_term", None) is not None) else None
            C = int(local[0]["correspondences"].shape[1])
            if old_corr is not None and old_corr.shape[1] != C:
                old_corr = old_corr[:, :C]       # (pad_records appended a zero-weight record to an odd C)
            if old_corr is None and not local[0]["correspondences"].is_cuda:
                # the largest upload, needed by one kernel only: it goes LAST on the upload stream, and nothing on the
                # compute stream waits for it until the first iteration's correspondence kernel (FusedJointOpt._run)
                cur, up = torch.cuda.current_stream(), upload_stream(K.device, deferred=True)
                up.wait_stream(cur)
                with torch.cuda.stream(up):
                    rec = frames_on_device(sh, "correspondences", pre_corr)
                    rec.record_stream(cur)
                    corr = CorrespondenceTerm(rec, K, image_size=int(masks.shape[-1]), delta=corr_delta)
                    corr.records.record_stream(cur)
                    corr.w_local.record_stream(cur)
                    corr.ready = up.record_event()
            else:
                corr = frames_on_device(sh, "correspondences", (old_corr, s0))
        return Joint_Optimizer(
            translations_object=trans, rotations_object=rots, verts_object_og=verts_object_og,
            faces_object=faces_dev, target_masks_object=masks, camintr_rois_object=K,
            int_scale_init=1, optimize_object_scale=optimize_object_scale, correspondences=corr,
            corr_delta=corr_delta)

    mark("mesh")
    model, cost = None, None
    if wants_probe(balance, num_iterations, shard.world) and auto:
        # cost-weighted partition: time the heavy kernels block by block on the equal-count ranges, gather, re-cut
        nblocks = max(1, min(16, (B_total // shard.world) // 32))
        model0 = build(shard, with_corr=False)
        pre = start_corr_upload(shard)      # runs beside the probe; the re-cut range reuses what it has
        lw_probe = {k: v for k, v in loss_weights.items() if k != "lw_corr_obj"}
        with FusedJointOpt(model0, lw_probe, lr, 0, shard=shard, keep_sum=1.0, exchange=False) as probe:
            ms = torch.from_numpy(probe.probe(nblocks)).cuda()
        ms_all = allgather_equal(ms, shard).cpu().numpy()
        cost = np.concatenate([frame_costs_from_blocks(ms_all[r, :nblocks], shard.bounds[r], shard.bounds[r + 1])
                               for r in range(shard.world)])
        new = shard.with_bounds(balanced_bounds(cost, shard.world))
        model = model0 if (new.start, new.stop) == (shard.start, shard.stop) and not corr_on else \
            build(new, with_corr=True, reuse=(model0, shard), pre_corr=pre)
        shard = new
        del model0
    mark("partition")
    if model is None:
        model = build(shard, with_corr=True)
    mark("model")
    fused = FusedJointOpt(model, loss_weights, lr, num_iterations, shard=shard, halo=halo, corr_on=corr_on)
    mark("fused")

    def rebalance(fused, model, shard, done):
        """Second cut, from the run itself: every rank's own time per iteration (the kernels' device clock, waits on
        the neighbours excluded) over the last iterations rescales the probe's per-frame costs inside its range; if the
        slowest rank is more than 3 % above the mean the ranges are re-cut and the moved frames change hands -- poses
        and Adam moments through one all-gather, masks / correspondences from the host list (only the frames a rank
        did not hold).  The iteration counter, history rows and the shared scale carry over."""
        t = torch.tensor([float(np.mean(fused.compute_ms(max(done - 8, 0), done)))], dtype=torch.float64, device="cuda")
        t_all = allgather_equal(t, shard).reshape(-1).cpu().numpy()
        if _rebalance_to is None and not (t_all.max() > 1.03 * t_all.mean()):
            return fused, model, shard
        new = shard.with_bounds(_rebalance_to if _rebalance_to is not None else
                                balanced_bounds(rescale_costs(cost, shard.bounds, t_all), shard.world))
        if new.bounds == shard.bounds:
            return fused, model, shard
        B = shard.B
        state = torch.cat([model.rotations_object.detach().reshape(B, 6), model.translations_object.detach().reshape(B, 3),
                           fused.m_rot, fused.v_rot, fused.m_tr, fused.v_tr], 1)
        state = allgather_frames(state, shard)[new.start:new.stop]
        model2 = build(new, with_corr=True, reuse=(model, shard))
        with torch.no_grad():
            model2.rotations_object.copy_(state[:, 0:6].reshape(-1, 3, 2))
            model2.translations_object.copy_(state[:, 6:9].reshape(-1, 1, 3))
            model2.int_scales_object.copy_(model.int_scales_object)
        fused2 = FusedJointOpt(model2, loss_weights, lr, num_iterations, shard=new, halo=halo, corr_on=corr_on,
                               start_step=done)
        fused2.m_rot.copy_(state[:, 9:15])
        fused2.v_rot.copy_(state[:, 15:21])
        fused2.m_tr.copy_(state[:, 21:24])
        fused2.v_tr.copy_(state[:, 24:27])
        fused2.mv_scale.copy_(fused.mv_scale)
        fused2.hist[:done].copy_(fused.hist[:done])     # partial sums of the old range: the sum over ranks is unchanged
        fused.check_status()
        fused.release()
        return fused2, model2, new

    try:
        try:
            from tqdm.auto import tqdm
            loop = tqdm(total=num_iterations)
        except Exception:  # pragma: no cover
            loop = None
        can_rebalance = cost is not None and num_iterations >= 64
        chunk = 50
        done = 0
        while done < num_iterations:
            n = min(16 if (can_rebalance and done == 0) else chunk, num_iterations - done)
            fused.run(n, use_graph=use_graph)
            done += n
            if loop is not None:
                loop.update(n)
            if can_rebalance and done == 16:
                fused, model, shard = rebalance(fused, model, shard, done)
                mark("rebalance")
        mark("launched")
        faces.join()     # the face-list identity check that ran beside the launches; raises before anything is returned
        loss_evolution = fused.history()  # the only host synchronisation of the loop
        mark("loop")
        if loop is not None:
            if loss_evolution.get("loss"):
                loop.set_description(f"Loss {loss_evolution['loss'][-1]:.4f}")
            loop.close()
    finally:
        fused.release()
    if board is not None:
        for k in ("loss_smooth_obj", "loss_sil_obj", "loss_corr_obj"):
            for step, val in enumerate(loss_evolution.get(k, [])):
                board.add_scalar(k, val, step)
    if shard.world > 1:
        with torch.no_grad():
            pose = torch.cat([model.rotations_object.detach().reshape(shard.B, 6),
                              model.translations_object.detach().reshape(shard.B, 3)], 1)
            pose = allgather_frames(pose, shard)   # one collective for both parameter tensors
            model.rotations_object = nn.Parameter(pose[:, :6].reshape(-1, 3, 2).contiguous())
            model.translations_object = nn.Parameter(pose[:, 6:].reshape(-1, 1, 3).contiguous())
    mark("gather")
    model.frame_shard = shard
    model.timing = [(b[0], (b[1] - a[1]) * 1e3) for a, b in zip(marks, marks[1:])]
    return model, loss_evolution
