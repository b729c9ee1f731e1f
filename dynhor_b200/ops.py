"""torch.library registration of the C-ABI entry points: the "thin C-ABI torch custom op and autograd.Function"
BASELINE.json's north_star names.  ctypes (dynhor_b200/_lib.py) stays underneath; the ops add
schemas, fake-tensor (meta) implementations -- so that torch.compile / export / FakeTensorMode can trace through
code that calls the kernels -- and the autograd formula of the silhouette renderer.

    torch.ops.dynhor.sil_forward(verts[B,V,3], faces[F,3] i32, K[B,3,3], image_size, anti_aliasing, near, far, eps,
                                 orig_size) -> rend[B,S,S]                      utils/losses.py:36-40,68
    torch.ops.dynhor.sil_backward(verts, faces, K, grad_rend, ...same...) -> grad_verts[B,V,3]   autograd of :68
    torch.ops.dynhor.jointopt_run(rot6d!, trans!, scale!, handle, n_iters, use_graph)           jointopt.py:144-160
    torch.ops.dynhor.dino_topk(frame_bank[Fm,K] bf16, templ_bank[N,K] bf16, k) -> (scores[Fm,N], vals[Fm,k], idx[Fm,k])
                                                                                pose_initializtion.py:295-311
The real implementations need CUDA tensors and the native library (no CPU fallback); the fake ones need neither.
"""
import weakref
from collections import OrderedDict
from typing import Tuple

import torch
from torch.library import custom_op

from . import _lib

# ----------------------------------------------------------------------------------------- renderer state cache
# The kernels work in caller-owned scratch (include/dynhor_b200.h): one SilhouetteState per problem shape and input
# pair (faces, K), found again by the backward op.  Small LRU; an entry holds ~1.5 MB per frame.
_STATES = OrderedDict()
_MAX_STATES = 4


def silhouette_state(verts, faces, K, image_size, anti_aliasing, near, far, eps, orig_size):
    from .renderer import SilhouetteState
    B, V = int(verts.shape[0]), int(verts.shape[1])
    key = (verts.device, B, V, int(faces.shape[0]), int(image_size), bool(anti_aliasing), float(near), float(far),
           float(eps), float(orig_size), faces.data_ptr(), faces._version, K.data_ptr(), K._version)
    st = _STATES.get(key)
    if st is None:
        st = SilhouetteState(B, V, faces, K, int(image_size), bool(anti_aliasing), float(near), float(far), float(eps),
                             float(orig_size))
        _STATES[key] = st
        while len(_STATES) > _MAX_STATES:
            _STATES.popitem(last=False)
    else:
        _STATES.move_to_end(key)
    return st


def _need_cuda(*ts):
    for t in ts:
        if not t.is_cuda:
            raise _lib.DynhorError("dynhor ops need CUDA tensors (dynhor_b200 has no CPU fallback)")


# ----------------------------------------------------------------------------------------- silhouette renderer
@custom_op("dynhor::sil_forward", mutates_args=())
def sil_forward(verts: torch.Tensor, faces: torch.Tensor, K: torch.Tensor, image_size: int, anti_aliasing: bool,
                near: float, far: float, eps: float, orig_size: float) -> torch.Tensor:
    _need_cuda(verts, faces, K)
    st = silhouette_state(verts, faces, K, image_size, anti_aliasing, near, far, eps, orig_size)
    rend, v = st.forward(verts)
    st.last_verts = (v.data_ptr(), v._version, st.version)
    return rend


@sil_forward.register_fake
def _(verts, faces, K, image_size, anti_aliasing, near, far, eps, orig_size):
    torch._check(verts.dim() == 3 and verts.shape[-1] == 3, lambda: "vertices must be [B,V,3]")
    torch._check(faces.dim() == 2 and faces.shape[-1] == 3, lambda: "faces must be [F,3]")
    return verts.new_empty((verts.shape[0], image_size, image_size), dtype=torch.float32)


@custom_op("dynhor::sil_backward", mutates_args=())
def sil_backward(verts: torch.Tensor, faces: torch.Tensor, K: torch.Tensor, grad_rend: torch.Tensor, image_size: int,
                 anti_aliasing: bool, near: float, far: float, eps: float, orig_size: float) -> torch.Tensor:
    _need_cuda(verts, faces, K, grad_rend)
    st = silhouette_state(verts, faces, K, image_size, anti_aliasing, near, far, eps, orig_size)
    v = verts.detach().contiguous().float()
    if getattr(st, "last_verts", None) != (v.data_ptr(), v._version, st.version):
        _, v = st.forward(v)     # the state's maps belong to another forward: rebuild them for these vertices
        st.last_verts = (v.data_ptr(), v._version, st.version)
    return st.backward(v, grad_rend)


@sil_backward.register_fake
def _(verts, faces, K, grad_rend, image_size, anti_aliasing, near, far, eps, orig_size):
    return verts.new_empty(verts.shape, dtype=torch.float32)


def _sil_setup(ctx, inputs, output):
    verts, faces, K, image_size, anti_aliasing, near, far, eps, orig_size = inputs
    ctx.save_for_backward(verts, faces, K)
    ctx.args = (image_size, anti_aliasing, near, far, eps, orig_size)


def _sil_backward(ctx, grad_rend):
    verts, faces, K = ctx.saved_tensors
    g = torch.ops.dynhor.sil_backward(verts, faces, K, grad_rend.contiguous(), *ctx.args)
    return g, None, None, None, None, None, None, None, None


sil_forward.register_autograd(_sil_backward, setup_context=_sil_setup)


# ----------------------------------------------------------------------------------------- fused optimisation loop
_FUSED = weakref.WeakValueDictionary()   # handle -> FusedJointOpt (the op carries plain ints and the mutated tensors)


def register_fused(fused):
    handle = id(fused)
    _FUSED[handle] = fused
    return handle


@custom_op("dynhor::jointopt_run", mutates_args=("rot6d", "trans", "scale"))
def jointopt_run(rot6d: torch.Tensor, trans: torch.Tensor, scale: torch.Tensor, handle: int, n_iters: int,
                 use_graph: bool) -> None:
    fused = _FUSED.get(handle)
    if fused is None:
        raise _lib.DynhorError("dynhor::jointopt_run: unknown handle (the FusedJointOpt was released)")
    if (rot6d.data_ptr(), trans.data_ptr()) != (fused.model.rotations_object.data_ptr(),
                                                fused.model.translations_object.data_ptr()):
        raise _lib.DynhorError("dynhor::jointopt_run: the tensors are not the parameters this plan was built on")
    fused._run(n_iters, use_graph)


@jointopt_run.register_fake
def _(rot6d, trans, scale, handle, n_iters, use_graph):
    return None


# ----------------------------------------------------------------------------------------- DINO matching
@custom_op("dynhor::dino_topk", mutates_args=())
def dino_topk(frame_bank: torch.Tensor, templ_bank: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    from .dino_match import _dino_cos_topk
    _need_cuda(frame_bank, templ_bank)
    return _dino_cos_topk(frame_bank, templ_bank, k)


@dino_topk.register_fake
def _(frame_bank, templ_bank, k):
    torch._check(frame_bank.dim() == 2 and templ_bank.dim() == 2, lambda: "banks must be [n, P*D]")
    torch._check(frame_bank.shape[1] == templ_bank.shape[1], lambda: "banks disagree on P*D")
    Fm, N = frame_bank.shape[0], templ_bank.shape[0]
    return (frame_bank.new_empty((Fm, N), dtype=torch.float32), frame_bank.new_empty((Fm, k), dtype=torch.float32),
            frame_bank.new_empty((Fm, k), dtype=torch.int64))
