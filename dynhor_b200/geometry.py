"""Host-side mirror of ObjTracker/utils/geometry.py (6D rotations) with the CUDA op underneath.

rot6d_to_matrix / matrix_to_rot6d keep the reference's names, argument meaning and output layout
(geometry.py:7-38): rot_6d [B,3,2] (or anything viewable as that) -> R [B,3,3] whose columns are b1, b2, b3.
"""
import ctypes

import torch

from . import _lib


class _Rot6dToMatrix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rot_6d):
        r6 = rot_6d.detach().reshape(-1, 3, 2).contiguous().float()
        if not r6.is_cuda:
            raise _lib.DynhorError("dynhor_b200.geometry.rot6d_to_matrix needs CUDA tensors (no CPU fallback)")
        R = torch.empty(r6.shape[0], 3, 3, device=r6.device, dtype=torch.float32)
        _lib.check(_lib.load().dh_rot6d_to_matrix(_lib.ptr(r6), _lib.ptr(R), r6.shape[0], _lib.stream_ptr()),
                   "dh_rot6d_to_matrix")
        ctx.save_for_backward(r6)
        return R

    @staticmethod
    def backward(ctx, gR):
        (r6,) = ctx.saved_tensors
        # analytic Gram-Schmidt backward through autograd on the same expression (tiny tensors)
        with torch.enable_grad():
            x = r6.detach().clone().requires_grad_(True)
            R = _rot6d_to_matrix_torch(x)
            (g,) = torch.autograd.grad(R, x, gR)
        return g.reshape(-1, 3, 2)


def _rot6d_to_matrix_torch(rot_6d):
    """geometry.py:19-25 in torch ops (used for the backward of the CUDA op)."""
    rot_6d = rot_6d.reshape(-1, 3, 2)
    a1, a2 = rot_6d[:, :, 0], rot_6d[:, :, 1]
    b1 = torch.nn.functional.normalize(a1)
    b2 = torch.nn.functional.normalize(a2 - torch.einsum("bi,bi->b", b1, a2).unsqueeze(-1) * b1)
    b3 = torch.linalg.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), dim=-1)


def rot6d_to_matrix(rot_6d):
    """geometry.py:7-25."""
    return _Rot6dToMatrix.apply(rot_6d)


def matrix_to_rot6d(rotmat):
    """geometry.py:28-38: the first two columns."""
    return rotmat.reshape(-1, 3, 3)[:, :, :2]


def rotation_angle_difference(R1, R2):
    """utils/camera.py:4-9: relative rotation angle in degrees."""
    R_rel = R1 @ R2.transpose(1, 2)
    tr = R_rel.diagonal(dim1=-2, dim2=-1).sum(-1)
    cos_theta = torch.clamp(0.5 * (tr - 1), -1.0, 1.0)
    return (180.0 / torch.pi) * torch.acos(cos_theta)
