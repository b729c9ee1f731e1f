"""[BUILDER-DEFINED] dense-correspondence reprojection term (include/dynhor_b200.h, struct dh_corr).

BASELINE.json's north_star puts "reprojection residuals of the DKM dense correspondences" on the hot path, but the
reference has no such code (README.md:43 only mentions a data folder of the unreleased
reconstruction stage).  The term is therefore defined here, wired in the reference's own style -- a loss named
`loss_corr_obj` weighted by `loss_weights["lw_corr_obj"]` (jointopt.py:147-150) -- and OFF unless the caller
passes correspondences AND a positive weight, so configs/custom_shoes.yaml reproduces the reference.

    records [B,C,6] f32 = X[3] canonical mesh-space point, t[2] target in ROI unit-image coordinates, w weight
    e = S * (K_roi (c.x/zc, c.y/zc, 1) - t),  c = (|s| X) R_b + T_b,  zc = c.z + 1e-9          [ROI pixels]
    loss_corr_obj = sum w * huber_delta(|e|) / sum w

Parity: against oracle/corr_oracle.py (torch CPU autograd) only -- "parity unpinned" by construction.
"""
import ctypes

import torch

from . import _lib
from .constants import REND_SIZE


def pad_records(records):
    """[B,C,6] -> contiguous fp32 with an even C (one zero-weight record appended if needed: the kernel's bulk
    copies move 16-byte multiples)."""
    r = records.detach().float()
    assert r.ndim == 3 and r.shape[-1] == 6, "correspondences must be [B,C,6]"
    if r.shape[1] % 2:
        r = torch.cat([r, torch.zeros_like(r[:, :1])], 1)
    return r.contiguous()


def plan(B, C, sm_count=0):
    out = (ctypes.c_int32 * 3)()
    _lib.check(_lib.load().dh_corr_plan(int(B), int(C), int(sm_count), out), "dh_corr_plan")
    return {"grid": out[0], "nslots": out[1], "tiles_per_frame": out[2]}


class _CorrSums(torch.autograd.Function):
    """Per-frame sums of w * huber(|e|) with the pose gradient from the same streaming pass (dh_corr_eval)."""

    @staticmethod
    def forward(ctx, rotations, translations, scale_abs, records, K, S, delta):
        if not rotations.is_cuda:
            raise _lib.DynhorError("correspondence term needs CUDA tensors (no CPU fallback)")
        B, C = records.shape[:2]
        nslots = plan(B, C)["nslots"]
        R = rotations.detach().contiguous().float()
        T = translations.detach().reshape(B, 3).contiguous().float()
        s = scale_abs.detach().reshape(1).contiguous().float()
        Kc = K.detach().contiguous().float()
        part = torch.empty(B, nslots, 16, device=R.device, dtype=torch.float32)
        _lib.check(_lib.load().dh_corr_eval(_lib.ptr(records), B, C, _lib.ptr(R), _lib.ptr(T), _lib.ptr(s),
                                            _lib.ptr(Kc), int(S), float(delta), _lib.ptr(part), nslots,
                                            _lib.stream_ptr()), "dh_corr_eval")
        sums = part.double().sum(1)                       # [B,16], slots added in order
        ctx.save_for_backward(sums, R, s)
        ctx.t_shape = translations.shape
        return sums[:, 12].float()

    @staticmethod
    def backward(ctx, grad_out):
        sums, R, s = ctx.saved_tensors
        go = grad_out.double().reshape(-1, 1)
        gT = (go * sums[:, 0:3]).float().reshape(ctx.t_shape)
        XG = sums[:, 3:12].reshape(-1, 3, 3)
        gR = (go.reshape(-1, 1, 1) * s.double() * XG).float()
        gs = (go.reshape(-1) * (R.double() * XG).sum((1, 2))).sum().float().reshape(1)
        return gR, gT, gs, None, None, None, None


class CorrespondenceTerm:
    """Holds the records of the local frames and evaluates loss_corr_obj (composable autograd path)."""

    def __init__(self, records, K_roi, image_size=REND_SIZE, delta=1.0, w_sum=None, ready=None):
        """ready: a CUDA event after which `records` is valid -- the term was built on a side stream that is still
        uploading (joint_optimize); every use from another stream waits for it first."""
        self.records = pad_records(records)
        if not self.records.is_cuda:
            if not torch.cuda.is_available():
                raise _lib.DynhorError("correspondence term needs a CUDA device (no CPU fallback)")
            self.records = self.records.cuda()
        self.K = K_roi
        self.S = int(image_size)
        self.delta = float(delta)
        self.w_local = self.records[..., 5].double().sum()      # device scalar: no host synchronisation here
        self._w_sum = None if w_sum is None else float(w_sum)
        self.ready = ready

    def wait_ready(self):
        """Make the current stream wait for the records (no-op once done)."""
        if self.ready is not None:
            torch.cuda.current_stream().wait_event(self.ready)
            self.ready = None

    @property
    def w_sum(self):
        if self._w_sum is None:
            self.wait_ready()
            self._w_sum = float(self.w_local.item())
        return self._w_sum

    @w_sum.setter
    def w_sum(self, v):
        self._w_sum = float(v)

    def loss(self, rotations, translations, scale_abs):
        self.wait_ready()
        sums = _CorrSums.apply(rotations, translations, scale_abs, self.records, self.K, self.S, self.delta)
        return sums.sum() / self.w_sum
