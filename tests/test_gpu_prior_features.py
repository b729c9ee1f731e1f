"""GPU: the batched template-bank build (dynhor_b200/prior_features.py) against the view-by-view
restatement of pose_initializtion.py:188-246 (oracle/prior_features_oracle.py over torchvision's roi_align): crops
bit-exact, ROI intrinsics, feature masks, and the bf16 bank giving the oracle's scores / best views.  DINOv2 itself is
library code outside the path (dino.py loads it from torch.hub: no network here): a small deterministic patch
embedder stands in for it on both sides."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class TinyDino(torch.nn.Module):
    """ObjTracker/dino.py's interface (extract_features, smaller_edge_size, feat_size) over a fixed 14x14 patch conv."""

    def __init__(self, dim=32, edge=14 * 9):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self.proj = torch.nn.Conv2d(3, dim, 14, 14, bias=True)
        with torch.no_grad():
            self.proj.weight.copy_(torch.randn(self.proj.weight.shape, generator=g) * 0.05)
            self.proj.bias.copy_(torch.randn(dim, generator=g) * 0.1)
        self.smaller_edge_size, self.feat_size = edge, edge // 14

    def extract_features(self, x):
        return torch.tanh(self.proj(x)).flatten(2).transpose(1, 2)


def _views(n, H=384, W=384, seed=0):
    """Rendered-template stand-ins: RGBA float images with a rotated ellipse (alpha 1 inside), a depth map."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    rend = np.zeros((n, H, W, 4), np.float32)
    depth = np.zeros((n, H, W, 1), np.float32)
    for i in range(n):
        cy, cx = (rng.uniform(0.3 * H, 0.7 * H), rng.uniform(0.3 * W, 0.7 * W)) if i % 3 else (6.0, W - 7.0)
        a, b, th = rng.uniform(0.05 * H, 0.25 * H), rng.uniform(0.05 * W, 0.25 * W), rng.uniform(0, np.pi)
        u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
        v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
        m = (u / b) ** 2 + (v / a) ** 2 < 1
        rend[i, ..., :3] = rng.random((H, W, 3)).astype(np.float32) * m[..., None]
        rend[i, ..., 3] = m
        depth[i, ..., 0] = (2.0 + 0.3 * np.sin(xx / 17.0)) * m
    K = np.tile(np.array([[500.0, 0, W / 2], [0, 500.0, H / 2], [0, 0, 1]], np.float32), (n, 1, 1))
    return {"prior_batched_renderings": torch.from_numpy(rend), "prior_depths": torch.from_numpy(depth),
            "Ks": torch.from_numpy(K), "Rs": torch.eye(3)[None].repeat(n, 1, 1), "Ts": torch.zeros(n, 3)}


def test_template_bank_vs_view_by_view_oracle():
    from dynhor_b200.dino_match import build_bank, dino_cos_topk
    from dynhor_b200.prior_features import compute_prior_features
    from oracle import dino_oracle, prior_features_oracle
    n = 9
    prior = _views(n)
    dino = TinyDino()
    ref = prior_features_oracle.compute_prior_features(prior, dino)
    out = compute_prior_features(prior, dino.cuda(), batch_size=4, keep_fp32=True)      # ragged last batch
    assert torch.equal(out["render_crop_masks"].cpu(), ref["render_crop_masks"])
    assert torch.equal(out["render_crop_imgs"].cpu(), ref["render_crop_imgs"])           # ROIAlign: bit-exact
    assert torch.equal(out["render_crop_depths"].cpu(), ref["render_crop_depths"])
    assert torch.allclose(out["render_roi_Ks"].cpu(), ref["render_roi_Ks"], rtol=1e-6, atol=1e-4)
    assert torch.equal(out["render_feats_masks"].cpu(), ref["render_feats_masks"])
    assert torch.allclose(out["render_feats"].cpu(), ref["render_feats"], atol=2e-5)      # conv on GPU vs CPU
    assert out["templ_bank"].shape == (n, ref["render_feats"].shape[1] * ref["render_feats"].shape[2])
    assert out["templ_bank"].dtype == torch.bfloat16 and out["templ_bank"].is_cuda
    # the bank scores like the reference expression on the oracle's fp32 features
    P, D = ref["render_feats"].shape[1:]
    g = torch.Generator().manual_seed(1)
    frames = torch.nn.functional.normalize(ref["render_feats"][[2, 5, 7]] + 0.2 * torch.randn(3, P, D, generator=g), dim=-1)
    fmask = (torch.rand(3, P, generator=g) < 0.6).float()
    fmask[:, 0] = 1
    s_o, _, i_o = dino_oracle.dino_cos_topk(frames, fmask, ref["render_feats"], 3)
    s, _, i = dino_cos_topk(build_bank(frames.cuda(), fmask.cuda()), out["templ_bank"], 3)
    assert torch.allclose(s.cpu(), s_o, atol=4e-3) and torch.equal(i[:, 0].cpu(), torch.tensor([2, 5, 7]))
    assert torch.equal(i_o[:, 0], torch.tensor([2, 5, 7]))


def test_empty_view_raises_like_the_reference():
    from dynhor_b200.prior_features import crop_views
    prior = _views(2)
    prior["prior_batched_renderings"][1, ..., 3] = 0
    with pytest.raises(RuntimeError):
        crop_views(prior["prior_batched_renderings"], prior["prior_depths"])
