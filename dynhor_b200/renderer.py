"""Silhouette renderer behind the `neural_renderer` call signature the reference uses.

Replaces, for camera_mode="projection" and mode="silhouettes":
    nr.renderer.Renderer(image_size, K, R, t, orig_size[, anti_aliasing])      utils/losses.py:36-40
    renderer(verts, faces, mode="silhouettes") -> [B, image_size, image_size]  utils/losses.py:68
    nr.projection(verts, K, R, t, dist_coeffs, orig_size)                      utils/losses.py:48-55
The rasteriser forward and its pseudo-gradient backward are the CUDA kernels of libdynhor_b200.so
(dh_sil_forward / dh_sil_backward), reached through the torch.library ops dynhor::sil_forward / dynhor::sil_backward
(dynhor_b200/ops.py); everything fails loudly without them.
"""
import ctypes

import torch

from . import _lib

DEFAULT_NEAR = 0.1
DEFAULT_FAR = 100.0
DEFAULT_EPS = 1e-4


def projection(vertices, K, R, t, dist_coeffs, orig_size, eps=1e-9):
    """utils/camera.py:26-63 in torch ops (used by the off-screen penalty, not by the rasteriser kernels)."""
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


def shared_faces(faces):
    """[B,F,3] (identical copies, run.py:158) or [F,3] -> contiguous int32 [F,3] on the same device."""
    if faces.ndim == 3:
        if faces.shape[0] > 1 and not bool((faces == faces[:1]).all()):
            raise NotImplementedError(
                "dynhor_b200 renders one mesh topology for all frames (run.py:158 stacks identical faces); "
                "per-frame face lists are not supported")
        faces = faces[0]
    if faces.ndim != 2 or faces.shape[-1] != 3:
        raise AssertionError("Invalid shape for faces")
    return faces.to(torch.int32).contiguous()


class SilhouetteState:
    """Device buffers + the dh_sil descriptor for one (B, V, F, S, aa) problem."""

    def __init__(self, B, V, faces_i32, K, S, aa, near=DEFAULT_NEAR, far=DEFAULT_FAR, eps=DEFAULT_EPS,
                 orig_size=1.0):
        if not faces_i32.is_cuda:
            raise _lib.DynhorError("dynhor_b200 renderer needs CUDA tensors (no CPU fallback)")
        lib = _lib.load()
        dev = faces_i32.device
        self.B, self.V, self.F, self.S, self.aa = int(B), int(V), int(faces_i32.shape[0]), int(S), int(bool(aa))
        if int(faces_i32.min()) < 0 or int(faces_i32.max()) >= V:
            raise ValueError("face indices out of range")
        self.faces = faces_i32
        self.K = K.detach().reshape(-1, 3, 3).expand(B, 3, 3).contiguous().float()
        sizes = (ctypes.c_int64 * 13)()
        _lib.check(lib.dh_sil_scratch_bytes(self.B, self.V, self.F, self.S, self.aa, sizes), "dh_sil_scratch_bytes")
        self.buffers = [torch.empty(int(n), dtype=torch.uint8, device=dev) for n in sizes]
        c = _lib.DhSil()
        c.B, c.V, c.F, c.S, c.aa = self.B, self.V, self.F, self.S, self.aa
        c.near_, c.far_, c.eps, c.orig_size = float(near), float(far), float(eps), float(orig_size)
        c.faces, c.K = self.faces.data_ptr(), self.K.data_ptr()
        (c.proj, c.bin_count, c.bins, c.fidx, c.alpha_bits, c.pos_pool, c.neg_pool, c.gpool, c.gmax,
         c.owned, c.negT, c.row_rng, c.neg_lists) = [
            b.data_ptr() for b in self.buffers]
        self.c = c
        self.version = 0
        self.image_size = self.S * 2 if self.aa else self.S

    def face_index_map(self):
        """[B,is,is] int32, rasteriser row order (before the vertical flip)."""
        n = self.image_size
        return self.buffers[3].view(torch.int32).view(self.B, n, n)

    def coverage_bits(self):
        """[B,is,is/32] uint32-as-int32 bitmaps, bit i of word w = pixel 32*w+i."""
        n = self.image_size
        return self.buffers[4].view(torch.int32).view(self.B, n, n // 32)

    def forward(self, verts_cam):
        v = verts_cam.detach().contiguous().float()
        if v.shape != (self.B, self.V, 3):
            raise AssertionError(f"vertices must be [{self.B},{self.V},3], got {tuple(v.shape)}")
        rend = torch.empty(self.B, self.S, self.S, device=v.device, dtype=torch.float32)
        _lib.check(_lib.load().dh_sil_forward(ctypes.byref(self.c), _lib.ptr(v), _lib.ptr(rend), _lib.stream_ptr()),
                   "dh_sil_forward")
        self.version += 1
        return rend, v

    def backward(self, verts_cam, grad_rend):
        g = grad_rend.detach().contiguous().float()
        gv = torch.empty(self.B, self.V, 3, device=g.device, dtype=torch.float32)
        _lib.check(_lib.load().dh_sil_backward(ctypes.byref(self.c), _lib.ptr(verts_cam), _lib.ptr(g), _lib.ptr(gv),
                                               _lib.stream_ptr()), "dh_sil_backward")
        return gv


class Renderer(torch.nn.Module):
    """`neural_renderer.Renderer` for the silhouette mode the reference uses (losses.py:36-40,68)."""

    def __init__(self, image_size=256, anti_aliasing=True, background_color=(0, 0, 0), fill_back=True,
                 camera_mode="projection", K=None, R=None, t=None, dist_coeffs=None, orig_size=1024,
                 near=DEFAULT_NEAR, far=DEFAULT_FAR, **_unused):
        super().__init__()
        if camera_mode != "projection":
            raise NotImplementedError("dynhor_b200 Renderer implements camera_mode='projection' only")
        if not fill_back:
            raise NotImplementedError("dynhor_b200 Renderer implements fill_back=True only (the reference default)")
        self.image_size = image_size
        self.anti_aliasing = anti_aliasing
        self.background_color = background_color
        self.fill_back = fill_back
        self.camera_mode = camera_mode
        self.K, self.R, self.t = K, R, t
        if dist_coeffs is None and K is not None:
            dist_coeffs = torch.zeros(1, 5, device=K.device)
        self.dist_coeffs = dist_coeffs
        self.orig_size = orig_size
        self.near, self.far = near, far
        self.rasterizer_eps = DEFAULT_EPS
        self._state = None
        self._faces_key, self._faces_i32 = None, None

    def _faces(self, faces):
        """[B,F,3] / [F,3] -> the shared int32 [F,3] list, checked once per faces tensor (not per call)."""
        key = (faces.data_ptr(), faces._version, tuple(faces.shape))
        if self._faces_key != key:
            self._faces_i32, self._faces_key = shared_faces(faces), key
        return self._faces_i32

    def forward(self, vertices, faces, textures=None, mode=None, K=None, R=None, t=None, dist_coeffs=None,
                orig_size=None):
        if mode != "silhouettes":
            raise NotImplementedError("dynhor_b200 Renderer implements mode='silhouettes' only")
        return self.render_silhouettes(vertices, faces, K, R, t, dist_coeffs, orig_size)

    def render_silhouettes(self, vertices, faces, K=None, R=None, t=None, dist_coeffs=None, orig_size=None):
        K = self.K if K is None else K
        R = self.R if R is None else R
        t = self.t if t is None else t
        dist_coeffs = self.dist_coeffs if dist_coeffs is None else dist_coeffs
        if orig_size is not None and float(orig_size) != float(self.orig_size):
            raise NotImplementedError("per-call orig_size override")
        if dist_coeffs is not None and bool((dist_coeffs != 0).any()):
            raise NotImplementedError("lens distortion coefficients must be zero (the reference never sets them)")
        if K is None:
            raise ValueError("Renderer needs K (camera_mode='projection')")
        # extrinsics: the reference always passes R = I, t = 0 (losses.py:38-39); anything else is applied here
        if R is not None and t is not None:
            eye = torch.eye(3, device=R.device, dtype=R.dtype).expand_as(R)
            if not (bool((R == eye).all()) and bool((t == 0).all())):
                vertices = torch.matmul(vertices, R.transpose(2, 1)) + t
        if not vertices.is_cuda:
            raise _lib.DynhorError("dynhor_b200 renderer needs CUDA tensors (no CPU fallback)")
        from . import ops   # registers torch.ops.dynhor.*
        f = self._faces(faces)
        args = (int(self.image_size), bool(self.anti_aliasing), float(self.near), float(self.far),
                float(self.rasterizer_eps), float(self.orig_size))
        # torch.library custom op: C-ABI forward, edge-scan pseudo-gradient registered as its autograd formula
        rend = torch.ops.dynhor.sil_forward(vertices, f, K, *args)
        self._state = ops.silhouette_state(vertices, f, K, *args)   # the scratch maps of this call (tests, debugging)
        return rend
