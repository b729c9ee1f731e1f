"""oracle/nr_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle; never a product path).

A CPU stand-in for the third-party `neural_renderer` package as the reference uses it:
  * constructor  ObjTracker/utils/losses.py:36-40, ObjTracker/pose_initializtion.py:98-105
  * call         ObjTracker/utils/losses.py:68,   ObjTracker/pose_initializtion.py:146-147,160
  * projection   ObjTracker/utils/losses.py:48-55 (semantics: vendored copy ObjTracker/utils/camera.py:26-63)

The package itself (requirements.txt:8, hassony2/multiperson@master, unpinned) is NOT under /root/reference;
its published pipeline for mode="silhouettes" is restated here:
  fill_back -> projection -> vertices_to_faces -> rasterise at 2x (anti_aliasing) -> alpha -> vertical flip
  -> 2x2 average pool.
Rasteriser arithmetic: oracle/nmr_oracle.c through ctypes.  PARITY UNPINNED for the rasteriser internals;
the projection is pinned against ObjTracker/utils/camera.py:26-63 by tests/golden/make_golden.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes
import os
import subprocess
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DEFAULT_NEAR = 0.1
DEFAULT_FAR = 100.0
DEFAULT_EPS = 1e-4


def build_lib(force=False):
    """Compile oracle/nmr_oracle.c -> oracle/libnmr_oracle.so (gcc, OpenMP)."""
    so = os.path.join(_HERE, "libnmr_oracle.so")
    src = os.path.join(_HERE, "nmr_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libnmr_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build_lib())
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        L.nmr_forward.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                  ip, fp, fp, fp]
        L.nmr_forward.restype = None
        L.nmr_backward.argtypes = [fp, ip, fp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float]
        L.nmr_backward.restype = None
        _LIB = L
    return _LIB


def _fptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _iptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def rasterize_forward_np(face_verts, image_size, near=DEFAULT_NEAR, far=DEFAULT_FAR):
    """face_verts [B,NF,3,3] f32 (NDC x,y + depth) -> dict of un-flipped maps at `image_size`."""
    fv = np.ascontiguousarray(face_verts, dtype=np.float32)
    B, NF = fv.shape[:2]
    s = int(image_size)
    fidx = np.empty((B, s, s), np.int32)
    wmap = np.empty((B, s, s, 3), np.float32)
    dmap = np.empty((B, s, s), np.float32)
    amap = np.empty((B, s, s), np.float32)
    lib().nmr_forward(_fptr(fv), B, NF, s, near, far, _iptr(fidx), _fptr(wmap), _fptr(dmap), _fptr(amap))
    return {"face_index": fidx, "weight": wmap, "depth": dmap, "alpha": amap}


def rasterize_backward_np(face_verts, face_index, alpha, grad_alpha, eps=DEFAULT_EPS):
    fv = np.ascontiguousarray(face_verts, dtype=np.float32)
    B, NF = fv.shape[:2]
    s = face_index.shape[-1]
    fidx = np.ascontiguousarray(face_index, np.int32)
    a = np.ascontiguousarray(alpha, np.float32)
    g = np.ascontiguousarray(grad_alpha, np.float32)
    out = np.empty_like(fv)
    lib().nmr_backward(_fptr(fv), _iptr(fidx), _fptr(a), _fptr(g), _fptr(out), B, NF, s, eps)
    return out


class _RasterizeSilhouette(torch.autograd.Function):
    """alpha map of the rasteriser (un-flipped), with the edge-scan pseudo-gradient as backward."""

    @staticmethod
    def forward(ctx, face_verts, image_size, near, far, eps):
        fv = face_verts.detach().contiguous().float()
        maps = rasterize_forward_np(fv.numpy(), image_size, near, far)
        ctx.save_for_backward(fv)
        ctx.maps = maps
        ctx.eps = eps
        return torch.from_numpy(maps["alpha"])

    @staticmethod
    def backward(ctx, grad_alpha):
        (fv,) = ctx.saved_tensors
        m = ctx.maps
        g = rasterize_backward_np(fv.numpy(), m["face_index"], m["alpha"],
                                  grad_alpha.contiguous().float().numpy(), ctx.eps)
        return torch.from_numpy(g), None, None, None, None


def projection(vertices, K, R, t, dist_coeffs, orig_size, eps=1e-9):
    """Perspective projection to NDC, semantics of ObjTracker/utils/camera.py:26-63 (same op order)."""
    vertices = torch.matmul(vertices, R.transpose(2, 1)) + t
    x, y, z = vertices[:, :, 0], vertices[:, :, 1], vertices[:, :, 2]
    x_ = x / (z + eps)
    y_ = y / (z + eps)
    k1, k2, p1, p2, k3 = [dist_coeffs[:, None, i] for i in range(5)]
    r = torch.sqrt(x_ ** 2 + y_ ** 2)
    radial = 1 + k1 * (r ** 2) + k2 * (r ** 4) + k3 * (r ** 6)
    x__ = x_ * radial + 2 * p1 * x_ * y_ + p2 * (r ** 2 + 2 * x_ ** 2)
    y__ = y_ * radial + p1 * (r ** 2 + 2 * y_ ** 2) + 2 * p2 * x_ * y_
    vertices = torch.stack([x__, y__, torch.ones_like(z)], dim=-1)
    vertices = torch.matmul(vertices, K.transpose(1, 2))
    u, v = vertices[:, :, 0], vertices[:, :, 1]
    v = orig_size - v
    u = 2 * (u - orig_size / 2.) / orig_size
    v = 2 * (v - orig_size / 2.) / orig_size
    return torch.stack([u, v, z], dim=-1)


def vertices_to_faces(vertices, faces):
    """[B,V,3], [B,NF,3] int -> [B,NF,3,3] (gather)."""
    bs, nv = vertices.shape[:2]
    faces = faces.long() + (torch.arange(bs, dtype=torch.long, device=faces.device) * nv)[:, None, None]
    return vertices.reshape(bs * nv, 3)[faces]


def rasterize_silhouettes(face_verts, image_size=256, anti_aliasing=True, near=DEFAULT_NEAR, far=DEFAULT_FAR,
                          eps=DEFAULT_EPS):
    s = image_size * 2 if anti_aliasing else image_size
    alpha = _RasterizeSilhouette.apply(face_verts, s, near, far, eps)
    alpha = alpha.flip(1)  # vertical flip (row order reversed)
    if anti_aliasing:
        alpha = F.avg_pool2d(alpha[:, None, :, :], kernel_size=(2, 2))[:, 0]
    return alpha


class Renderer(torch.nn.Module):
    """mode="silhouettes", camera_mode="projection" subset of neural_renderer.Renderer."""

    def __init__(self, image_size=256, anti_aliasing=True, background_color=(0, 0, 0), fill_back=True,
                 camera_mode="projection", K=None, R=None, t=None, dist_coeffs=None, orig_size=1024,
                 near=DEFAULT_NEAR, far=DEFAULT_FAR, **_unused):
        super().__init__()
        if camera_mode != "projection":
            raise ValueError("oracle Renderer restates camera_mode='projection' only")
        self.image_size = image_size
        self.anti_aliasing = anti_aliasing
        self.background_color = background_color
        self.fill_back = fill_back
        self.camera_mode = camera_mode
        self.K, self.R, self.t = K, R, t
        if dist_coeffs is None:
            dist_coeffs = torch.zeros(1, 5)
        self.dist_coeffs = dist_coeffs
        self.orig_size = orig_size
        self.near, self.far = near, far
        self.rasterizer_eps = DEFAULT_EPS

    def forward(self, vertices, faces, textures=None, mode=None, K=None, R=None, t=None, dist_coeffs=None,
                orig_size=None):
        if mode != "silhouettes":
            raise ValueError("oracle Renderer restates mode='silhouettes' only")
        return self.render_silhouettes(vertices, faces, K, R, t, dist_coeffs, orig_size)

    def render_silhouettes(self, vertices, faces, K=None, R=None, t=None, dist_coeffs=None, orig_size=None):
        if self.fill_back:
            faces = torch.cat((faces, faces.flip(-1)), dim=1)
        K = self.K if K is None else K
        R = self.R if R is None else R
        t = self.t if t is None else t
        dist_coeffs = self.dist_coeffs if dist_coeffs is None else dist_coeffs
        orig_size = self.orig_size if orig_size is None else orig_size
        vertices = projection(vertices, K, R, t, dist_coeffs, orig_size)
        face_verts = vertices_to_faces(vertices, faces)
        return rasterize_silhouettes(face_verts, self.image_size, self.anti_aliasing, self.near, self.far,
                                     self.rasterizer_eps)


def as_neural_renderer_module():
    """A module object shaped like `neural_renderer` (nr.renderer.Renderer, nr.projection, ...), so the
    reference's own Python (utils/losses.py, jointopt.py) can be run on CPU by tests/golden/make_golden.py."""
    nr = types.ModuleType("neural_renderer")
    nr.renderer = types.ModuleType("neural_renderer.renderer")
    nr.renderer.Renderer = Renderer
    nr.Renderer = Renderer
    nr.projection = projection
    nr.vertices_to_faces = vertices_to_faces
    nr.rasterize_silhouettes = rasterize_silhouettes
    return nr


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv))
