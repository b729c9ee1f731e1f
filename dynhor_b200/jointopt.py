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
from collections import defaultdict

import torch
from torch import nn

from . import _lib
from .camera import compute_transformation_persp, tensorify
from .corr import CorrespondenceTerm, plan as corr_plan
from .geometry import matrix_to_rot6d, rot6d_to_matrix
from .losses import Losses
from .renderer import SilhouetteState, shared_faces
from .sharding import FrameShard, allgather_frames, allreduce_sum_, detect_shard, exchange_halo


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
        if correspondences is not None:
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
                 exchange=True, halo="p2p"):
        """halo: how boundary poses travel between ranks when the sequence is sharded.  "p2p" (default): CUDA-IPC
        mailboxes written by the update kernel itself over NVLink, no host work per iteration.  "nccl": one grouped
        send/recv per iteration driven from the host (also the path the gloo CPU tests exercise)."""
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
        self.sil = SilhouetteState(B, V, faces, model.camintr_rois_object, S, True, orig_size=1.0)
        self.verts = verts
        st = _lib.stream_ptr()
        # static inputs
        self.mask_tri = torch.empty(B, S, S, dtype=torch.int8, device=dev)
        keep = torch.zeros(1, dtype=torch.int64, device=dev)
        _lib.check(lib.dh_masks_prepare(_lib.ptr(masks.contiguous().float()), _lib.ptr(self.mask_tri), _lib.ptr(keep),
                                        masks.numel(), st), "dh_masks_prepare")
        self.exchange = exchange  # False: the caller fills self.halo by hand (single-process shard emulation)
        if keep_sum is None:
            keep_f = keep.to(torch.float64)
            allreduce_sum_(keep_f, self.shard, group)
            keep_sum = float(keep_f.item())
        self.keep_sum = float(keep_sum)
        self.keep_local = keep
        self.moments = torch.empty(12, dtype=torch.float64, device=dev)
        _lib.check(lib.dh_mesh_moments(_lib.ptr(verts), V, _lib.ptr(self.moments), st), "dh_mesh_moments")
        # optimiser state / history
        self.scale = model.int_scales_object
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)  # noqa: E731
        self.m_rot, self.v_rot, self.m_tr, self.v_tr, self.mv_scale = z(B, 6), z(B, 6), z(B, 3), z(B, 3), z(2)
        self.step = z(1, dt=torch.int32)
        self.max_iters = int(max_iters)
        self.hist = z(max(self.max_iters, 1), 4, dt=torch.float64)
        self.halo = z(2, 9)
        self.edge = z(2, 9)
        import os
        nchunks = nchunks or int(os.environ.get("DH_BWD_CHUNKS", "0"))  # tuning knob
        self.nchunks = int(nchunks) if nchunks else lib.dh_jointopt_default_chunks(B, self.sil.F)
        sizes = (ctypes.c_int64 * 5)()
        _lib.check(lib.dh_jointopt_scratch_bytes(B, self.nchunks, sizes), "dh_jointopt_scratch_bytes")
        self.scratch = [torch.zeros(int(n), dtype=torch.uint8, device=dev) for n in sizes]
        self.loss_weights = dict(loss_weights)
        p = _lib.DhJointOpt()
        p.sil = self.sil.c
        p.verts_og, p.mask_tri = verts.data_ptr(), self.mask_tri.data_ptr()
        p.rot6d, p.trans, p.scale = rot.data_ptr(), tr.data_ptr(), self.scale.data_ptr()
        p.adam_m_rot, p.adam_v_rot = self.m_rot.data_ptr(), self.v_rot.data_ptr()
        p.adam_m_trans, p.adam_v_trans = self.m_tr.data_ptr(), self.v_tr.data_ptr()
        p.adam_mv_scale = self.mv_scale.data_ptr()
        p.step, p.hist, p.max_iters = self.step.data_ptr(), self.hist.data_ptr(), self.hist.shape[0]
        p.halo_prev = self.halo[0].data_ptr() if self.shard.has_prev else None
        p.halo_next = self.halo[1].data_ptr() if self.shard.has_next else None
        p.B_total = self.shard.B_total
        p.keep_sum = self.keep_sum
        p.lw_sil = float(loss_weights.get("lw_sil_obj", 0.0))
        p.lw_smooth = float(loss_weights.get("lw_smooth_obj", 0.0))
        p.lr = float(lr)
        p.optimize_scale = int(bool(getattr(model, "optimize_object_scale", False)))
        if p.optimize_scale and self.shard.world > 1:
            raise NotImplementedError("optimize_object_scale with frame sharding (needs a scale-gradient all-reduce)")
        p.moments = self.moments.data_ptr()
        (p.Rmat, p.smooth_terms, p.loss_counts, p.partials, p.frame_terms) = [b.data_ptr() for b in self.scratch]
        p.nchunks = self.nchunks
        # builder-defined correspondence term: on only with records AND a positive weight
        lw_corr = float(loss_weights.get("lw_corr_obj", 0.0))
        self.corr_on = getattr(model, "corr_term", None) is not None and lw_corr > 0
        if self.corr_on:
            ct = model.corr_term
            w = ct.w_local.reshape(1).clone()
            if exchange:
                allreduce_sum_(w, self.shard, group)
                self.corr_w_sum = float(w.item())
            else:
                self.corr_w_sum = ct.w_sum
            cp = corr_plan(B, ct.records.shape[1])
            self.corr_partials = z(B, cp["nslots"], 16)
            p.corr.records, p.corr.C, p.corr.nslots = ct.records.data_ptr(), ct.records.shape[1], cp["nslots"]
            p.corr.delta, p.corr.w_sum, p.corr.lw_corr = ct.delta, self.corr_w_sum, lw_corr
            p.corr.partials = self.corr_partials.data_ptr()
        self.p = p
        self.halo_mode = halo if (self.shard.world > 1 and exchange) else "none"
        self._mailbox, self._peers = None, []
        self._sync_halo()
        if self.halo_mode == "p2p":
            self._setup_p2p()

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
        """Allocate this rank's mailbox, swap CUDA-IPC handles with the neighbours and seed parity-0 slots with
        the initial halo (already exchanged once through the process group)."""
        import torch.distributed as dist
        lib = _lib.load()
        mb = ctypes.c_void_p()
        _lib.check(lib.dh_dev_alloc(ctypes.byref(mb), 4 * 128), "dh_dev_alloc")
        self._mailbox = mb
        handle = ctypes.create_string_buffer(64)
        _lib.check(lib.dh_ipc_export(mb, handle), "dh_ipc_export")
        handles = [None] * self.shard.world
        dist.all_gather_object(handles, bytes(handle.raw), group=self.group)
        st = _lib.stream_ptr()
        # slot(side, parity 0): side 0 = pose of the frame before our first, side 1 = after our last
        for side in (0, 1):
            _lib.check(lib.dh_memcpy_d2d(ctypes.c_void_p(mb.value + side * 2 * 16 * 4), _lib.ptr(self.halo[side]),
                                         9 * 4, st), "dh_memcpy_d2d")
        torch.cuda.synchronize()
        dist.barrier(group=self.group)  # every mailbox is seeded before anyone may publish into it
        peers = {}
        for name, r in (("peer_prev", self.shard.rank - 1), ("peer_next", self.shard.rank + 1)):
            if 0 <= r < self.shard.world:
                ptr = ctypes.c_void_p()
                _lib.check(lib.dh_ipc_open(ctypes.create_string_buffer(handles[r], 64), ctypes.byref(ptr)),
                           "dh_ipc_open")
                peers[name] = ptr
                self._peers.append(ptr)
        self.p.mailbox = mb.value
        self.p.peer_prev = peers["peer_prev"].value if "peer_prev" in peers else None
        self.p.peer_next = peers["peer_next"].value if "peer_next" in peers else None

    # -- execution --------------------------------------------------------------------------------------------
    def run(self, n_iters, use_graph=True):
        """n_iters fused iterations on the current stream; no host synchronisation (single rank)."""
        lib = _lib.load()
        if self.halo_mode != "nccl":
            _lib.check(lib.dh_jointopt_run(ctypes.byref(self.p), int(n_iters), int(use_graph), _lib.stream_ptr()),
                       "dh_jointopt_run")
            return
        for _ in range(int(n_iters)):
            _lib.check(lib.dh_jointopt_run(ctypes.byref(self.p), 1, int(use_graph), _lib.stream_ptr()),
                       "dh_jointopt_run")
            self._sync_halo()

    def grads(self):
        """Gradients of the weighted loss for the current parameters (no update)."""
        B = self.shard.B
        dev = self.step.device
        g_rot = torch.empty(B, 3, 2, device=dev)
        g_tr = torch.empty(B, 1, 3, device=dev)
        g_s = torch.zeros(1, device=dev)
        _lib.check(_lib.load().dh_jointopt_grads(ctypes.byref(self.p), _lib.ptr(g_rot), _lib.ptr(g_tr), _lib.ptr(g_s),
                                                 _lib.stream_ptr()), "dh_jointopt_grads")
        return g_rot, g_tr, g_s

    def evaluate(self):
        """Losses of the current parameters -> dict of floats (synchronises)."""
        _lib.check(_lib.load().dh_jointopt_eval(ctypes.byref(self.p), _lib.stream_ptr()), "dh_jointopt_eval")
        row = int(self.step.item())
        return self._rows_to_dict(self.hist[row:row + 1].clone())

    def _rows_to_dict(self, rows):
        rows = rows.clone()
        if self.exchange:
            allreduce_sum_(rows, self.shard, self.group)
        rows = rows.cpu().numpy()
        lw = self.loss_weights
        evo = defaultdict(list)
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
        n = int(self.step.item())
        return self._rows_to_dict(self.hist[:n])

    KERNELS = ("pose_prep", "project", "setup_bin", "raster", "backward", "pose_update", "finalize", "corr")

    def profile(self, n_iters=5):
        """Average per-kernel milliseconds over n_iters real iterations (CUDA events on the launch stream)."""
        ms = (ctypes.c_float * 8)()
        _lib.check(_lib.load().dh_jointopt_profile(ctypes.byref(self.p), int(n_iters), ms, _lib.stream_ptr()),
                   "dh_jointopt_profile")
        return {k: float(ms[i]) for i, k in enumerate(self.KERNELS)}

    def release(self):
        lib = _lib.load()
        lib.dh_jointopt_release(ctypes.byref(self.p))
        if self._mailbox is not None:
            import torch.distributed as dist
            torch.cuda.synchronize()
            dist.barrier(group=self.group)  # nobody may still be publishing into a mailbox that is about to go
            for ptr in self._peers:
                lib.dh_ipc_close(ptr)
            lib.dh_dev_free(self._mailbox)
            self._mailbox, self._peers = None, []
            self.p.mailbox = self.p.peer_prev = self.p.peer_next = None


def joint_optimize(object_parameters, objvertices=None, objfaces=None, loss_weights=None, num_iterations=400,
                   lr=1e-4, board=None, optimize_object_scale=False, shard=None, use_graph=True, halo="p2p",
                   corr_delta=1.0):
    """jointopt.py:93-161.  Extra keyword `shard` (a sharding.FrameShard): when given (or when torch.distributed
    is initialised with more than one rank) `object_parameters` is the full sequence and this rank optimises its
    contiguous frame range; the returned model holds the gathered poses of ALL frames on every rank."""
    if not torch.cuda.is_available():
        raise _lib.DynhorError("joint_optimize needs a CUDA device (dynhor_b200 has no CPU fallback)")
    if loss_weights is None:
        loss_weights = {"lw_sil_obj": 1.0, "lw_smooth_obj": 1.0}
    B_total = len(object_parameters)
    shard = detect_shard(B_total) if shard is None else shard
    verts_object_og = tensorify(objvertices).cuda()
    # faces: run.py:158 stacks one face list per frame; only this rank's frames are uploaded and checked
    faces_host = tensorify(objfaces)
    faces_local = (faces_host[shard.start:shard.stop] if faces_host.ndim == 3 else faces_host).cuda()
    local = shard.slice(object_parameters)
    # Stage 1 hands over CUDA tensors (pose_initializtion.py:460-471); host tensors are accepted too: the small
    # ones are concatenated on the host, the masks go up frame by frame (no 78 MB host-side concatenation) and the
    # ref / keep masks are derived on the device.
    obj_trans = torch.cat([obj["translations"] for obj in local]).cuda()
    obj_rots = torch.cat([obj["rotations"] for obj in local]).cuda()
    obj_camintr_roi = torch.cat([obj["K_roi"][:, 0] for obj in local]).cuda()
    obj_tar_masks = torch.cat([obj["target_masks"].cuda(non_blocking=True) for obj in local])
    # optional per-frame key "correspondences" [1,C,6] (builder-defined term, corr.py); absent -> reference behaviour
    corr = None
    if all("correspondences" in obj for obj in local) and loss_weights.get("lw_corr_obj", 0) > 0:
        corr = torch.cat([obj["correspondences"].cuda(non_blocking=True) for obj in local])
    model = Joint_Optimizer(
        translations_object=obj_trans, rotations_object=obj_rots, verts_object_og=verts_object_og,
        faces_object=faces_local, target_masks_object=obj_tar_masks, camintr_rois_object=obj_camintr_roi,
        int_scale_init=1, optimize_object_scale=optimize_object_scale, correspondences=corr,
        corr_delta=corr_delta)
    fused = FusedJointOpt(model, loss_weights, lr, num_iterations, shard=shard, halo=halo)
    try:
        from tqdm.auto import tqdm
        loop = tqdm(total=num_iterations)
    except Exception:  # pragma: no cover
        loop = None
    chunk = 50
    done = 0
    while done < num_iterations:
        n = min(chunk, num_iterations - done)
        fused.run(n, use_graph=use_graph)
        done += n
        if loop is not None:
            loop.update(n)
    loss_evolution = fused.history()  # the only host synchronisation of the loop
    if loop is not None:
        if loss_evolution.get("loss"):
            loop.set_description(f"Loss {loss_evolution['loss'][-1]:.4f}")
        loop.close()
    if board is not None:
        for k in ("loss_smooth_obj", "loss_sil_obj", "loss_corr_obj"):
            for step, val in enumerate(loss_evolution.get(k, [])):
                board.add_scalar(k, val, step)
    fused.release()
    if shard.world > 1:
        with torch.no_grad():
            model.rotations_object = nn.Parameter(allgather_frames(model.rotations_object.detach(), shard))
            model.translations_object = nn.Parameter(allgather_frames(model.translations_object.detach(), shard))
    return model, loss_evolution
