"""Host-side mirror of the pieces of ObjTracker/utils/camera.py on the joint-optimisation path."""
import torch

from . import _lib


def tensorify(array, device=None):
    """utils/camera.py:11-16."""
    if not isinstance(array, torch.Tensor):
        array = torch.tensor(array)
    if device is not None:
        array = array.to(device)
    return array


class _TransformVerts(torch.autograd.Function):
    """(|s| v) @ R + T  (utils/camera.py:204-206) as one kernel; analytic backward in torch ops."""

    @staticmethod
    def forward(ctx, verts, translations, rotations, scales):
        v = verts.detach().contiguous().float()
        T = translations.detach().reshape(-1, 3).contiguous().float()
        R = rotations.detach().reshape(-1, 3, 3).contiguous().float()
        s = scales.detach().reshape(-1)[:1].contiguous().float()
        B, V = R.shape[0], v.shape[0]
        out = torch.empty(B, V, 3, device=v.device, dtype=torch.float32)
        _lib.check(_lib.load().dh_transform_verts(_lib.ptr(v), _lib.ptr(R), _lib.ptr(T), _lib.ptr(s), _lib.ptr(out),
                                                  B, V, _lib.stream_ptr()), "dh_transform_verts")
        ctx.save_for_backward(v, R, s)
        return out

    @staticmethod
    def backward(ctx, g):
        v, R, s = ctx.saved_tensors
        sv = s.abs() * v                                       # [V,3]
        gT = g.sum(1, keepdim=True)                            # [B,1,3]
        gR = torch.einsum("vi,bvj->bij", sv, g)                # [B,3,3]
        gs = (torch.einsum("vi,bij,bvj->", v, R, g) * torch.sign(s)).reshape(1)
        return None, gT, gR, gs


def compute_transformation_persp(meshes, translations, rotations=None, intrinsic_scales=None):
    """utils/camera.py:179-207.  meshes [V,3] (shared mesh) or [B,V,3]; translations [B,1,3]; rotations [B,3,3];
    intrinsic_scales [1] or [B]."""
    B = translations.shape[0]
    device = meshes.device
    if rotations is None:
        rotations = torch.eye(3, device=device).unsqueeze(0).repeat(B, 1, 1)
    if intrinsic_scales is None:
        intrinsic_scales = torch.ones(1, device=device)
    if meshes.ndimension() == 2 and meshes.is_cuda and intrinsic_scales.numel() == 1:
        return _TransformVerts.apply(meshes, translations, rotations, intrinsic_scales)
    if meshes.ndimension() == 2:
        meshes = meshes.repeat(B, 1, 1)
    return torch.matmul(intrinsic_scales.view(-1, 1, 1) * meshes, rotations) + translations


def get_K_crop_resize(K, boxes, crop_resize, invert_xy=False):
    """Intrinsics of the image cropped to `boxes` [n,4] (x0, y0, x1, y1) and resized to `crop_resize`: the closed
    formula of utils/camera.py:84-130 (pixel centres at integer coordinates; no skew, like the reference), kept in
    the reference's floating-point operation order so that K_roi -- an input of the bit-exact rasteriser -- comes out
    identical.  Used for the ROI intrinsics of frames (pose_initializtion.py:275-277) and of template views (:217-219)."""
    assert K.shape[1:] == (3, 3)
    assert boxes.shape[1:] == (4,)
    K, boxes = K.float(), boxes.float()
    x0, y0, x1, y1 = (boxes[:, [1, 0, 3, 2]] if invert_xy else boxes).unbind(1)
    size = torch.tensor(crop_resize, dtype=torch.float)
    out = {"x": max(size), "y": min(size)}
    new_K = K.clone()
    for axis, lo, hi, row in (("x", x0, x1, 0), ("y", y0, y1, 1)):
        extent = hi - lo
        half = (extent - 1) / 2                                # the crop's centre pixel
        shifted = K[:, row, 2] + half - (lo + hi) / 2          # principal point in crop coordinates
        scale = out[axis] / extent
        new_K[:, row, row] = scale * K[:, row, row]
        new_K[:, row, 2] = (out[axis] - 1) / 2 + scale * (shifted - half)
    return new_K
