"""torch.library registration (dynhor_b200/ops.py): schemas and fake-tensor (meta) implementations, checked without a
GPU -- FakeTensorMode builds "cuda" tensors that own no memory, the fake kernels only compute shapes and dtypes.  The
real kernels behind the same ops are exercised by the -m gpu tests (every renderer / fused-loop / DINO call goes
through torch.ops.dynhor.*)."""
import pytest
import torch
from torch._subclasses.fake_tensor import FakeTensorMode

import dynhor_b200.ops  # noqa: F401  (registers the ops)


def test_schemas():
    ops = torch.ops.dynhor
    assert str(ops.sil_forward.default._schema) == (
        "dynhor::sil_forward(Tensor verts, Tensor faces, Tensor K, SymInt image_size, bool anti_aliasing, float near, "
        "float far, float eps, float orig_size) -> Tensor")
    s = str(ops.jointopt_run.default._schema)
    assert "Tensor(a0!) rot6d" in s and "Tensor(a1!) trans" in s and "Tensor(a2!) scale" in s and s.endswith("-> ()")
    assert str(ops.dino_topk.default._schema).endswith("-> (Tensor, Tensor, Tensor)")
    assert "grad_rend" in str(ops.sil_backward.default._schema)


def test_fake_tensor_shapes_and_dtypes():
    with FakeTensorMode():
        v = torch.empty(4, 100, 3, device="cuda", requires_grad=True)
        f = torch.empty(50, 3, dtype=torch.int32, device="cuda")
        K = torch.empty(4, 3, 3, device="cuda")
        r = torch.ops.dynhor.sil_forward(v, f, K, 256, True, 0.1, 100.0, 1e-4, 1.0)
        assert r.shape == (4, 256, 256) and r.dtype == torch.float32 and r.device.type == "cuda" and r.requires_grad
        g = torch.ops.dynhor.sil_backward(v.detach(), f, K, torch.empty_like(r), 256, True, 0.1, 100.0, 1e-4, 1.0)
        assert g.shape == (4, 100, 3) and g.dtype == torch.float32
        fb = torch.empty(300, 1369 * 8, dtype=torch.bfloat16, device="cuda")
        tb = torch.empty(1000, 1369 * 8, dtype=torch.bfloat16, device="cuda")
        scores, vals, idx = torch.ops.dynhor.dino_topk(fb, tb, 10)
        assert scores.shape == (300, 1000) and vals.shape == (300, 10) and idx.shape == (300, 10)
        assert idx.dtype == torch.int64 and scores.dtype == torch.float32
        assert torch.ops.dynhor.jointopt_run(torch.empty(4, 3, 2, device="cuda"), torch.empty(4, 1, 3, device="cuda"),
                                             torch.empty(1, device="cuda"), 0, 3, True) is None
        with pytest.raises(RuntimeError):
            torch.ops.dynhor.sil_forward(torch.empty(4, 100, 2, device="cuda"), f, K, 256, True, 0.1, 100.0, 1e-4, 1.0)
        with pytest.raises(RuntimeError):
            torch.ops.dynhor.dino_topk(fb, torch.empty(1000, 64, dtype=torch.bfloat16, device="cuda"), 10)


def test_real_kernels_refuse_cpu_tensors():
    """No CPU fallback behind the ops either."""
    from dynhor_b200._lib import DynhorError
    v, f, K = torch.zeros(1, 4, 3), torch.zeros(2, 3, dtype=torch.int32), torch.eye(3)[None]
    with pytest.raises(DynhorError):
        torch.ops.dynhor.sil_forward(v, f, K, 32, False, 0.1, 100.0, 1e-4, 1.0)
    with pytest.raises(DynhorError):
        torch.ops.dynhor.dino_topk(torch.zeros(2, 8, dtype=torch.bfloat16), torch.zeros(3, 8, dtype=torch.bfloat16), 1)
