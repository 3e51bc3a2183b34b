"""-m gpu: K12 fused loss kernels (forward sums and backward) against the torch restatement of the reference's
LapLoss / GradientLoss / weighted L1 (tests/ops_ref.py), fp32 on both sides.  Tolerance 2e-4 relative on the sums
(fp32 summation order), 2e-3 relative-L2 on gradients (sign flips of |.| at exact zeros are the only discontinuity)."""
import pytest
import torch

import ops_ref
from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("S,H,W", [(4, 64, 64), (6, 128, 96), (3, 32, 160), (20, 256, 256)])
def test_matte_loss_sums_forward_backward(S, H, W):
    from maggie_b200 import ops
    g = torch.Generator().manual_seed(S * H + W)
    tgt = synth.soft_ellipse_alphas(1, S, H, W, edge_px=5.0)[0].cuda()
    preds = [(tgt + 0.2 * torch.randn(S, H, W, generator=g).cuda()).clamp(0, 1) for _ in range(3)]
    ws = [(torch.rand(S, H, W, generator=g) > 0.6).float().cuda() for _ in range(2)] + \
         [(torch.rand(S, H, W, generator=g) > 0.5).float().cuda() + 1.0]
    ws[0][0] = 0  # an empty slice
    TERMS = ["weighted-L1", "lap level 0", "lap level 1", "lap level 2", "sobel", "all"]
    for term, name in enumerate(TERMS):
        gs = torch.zeros(3, 8).cuda()
        if name == "all":
            gs = torch.rand(3, 8, generator=g).cuda()
        else:
            gs[:, term] = torch.tensor([1.0, 0.5, 2.0]).cuda()
        _check(ops, S, H, W, preds, tgt, ws, gs, name)
    # per-plane factor on the predictions (the reference's `pred * valid_masks`), applied inside the kernels
    scale = (torch.rand(S, generator=g) > 0.3).float().cuda()
    _check(ops, S, H, W, preds, tgt, ws, torch.rand(3, 8, generator=g).cuda(), "all, plane scale", scale)


def _check(ops, S, H, W, preds, tgt, ws, gs, name, plane_scale=None):
    def run(fn):
        ps = [p.clone().requires_grad_(True) for p in preds]
        s = fn(ps[0], ps[1], ps[2], tgt, ws[0], ws[1], ws[2], plane_scale)
        (s * gs).sum().backward()
        return s.detach(), [p.grad for p in ps]

    s, gr = run(ops.matte_loss_sums)
    sr, grr = run(ops_ref.matte_loss_sums)
    assert s.shape == (3, 8)
    rel = ((s - sr).abs() / (sr.abs() + 1e-3)).max()
    assert float(rel) < 2e-4, (s, sr)
    for a, b in zip(gr, grr):
        assert a.shape == b.shape
        assert float((a - b).norm() / (b.norm() + 1e-12)) < 2e-3, f"{name}: rel L2 {float((a - b).norm() / (b.norm() + 1e-12))}"
        # borders exercise the reflect / replicate adjoints: check them separately
        for sl in (slice(0, 3), slice(-3, None)):
            assert float((a[:, sl] - b[:, sl]).abs().max()) < 2e-3 * float(b.abs().max()) + 1e-6
            assert float((a[:, :, sl] - b[:, :, sl]).abs().max()) < 2e-3 * float(b.abs().max()) + 1e-6


@pytest.mark.parametrize("S,H,W,band", [(6, 256, 256, 3.0), (4, 192, 320, 1.0), (3, 512, 512, 6.0)])
def test_matte_loss_with_band_weights_skipped_tiles(S, H, W, band):
    """The training configuration: the OS1 / OS4 weights are narrow bands (the refinement region), so most 32 x 32 tiles have
    no weight within 32 pixels and are skipped by the forward / backward passes (tile maps of K12): sums and gradients must
    still equal the dense evaluation, including tiles right next to the skipped ones."""
    from maggie_b200 import ops
    g = torch.Generator().manual_seed(S + H)
    tgt = synth.soft_ellipse_alphas(1, S, H, W, edge_px=band)[0].cuda()
    preds = [(tgt + 0.2 * torch.randn(S, H, W, generator=g).cuda()).clamp(0, 1) for _ in range(3)]
    edge = ((tgt > 0.02) & (tgt < 0.98)).float()
    ws = [edge.clone(), edge.clone(), torch.ones(S, H, W).cuda()]
    ws[1][0] = 0                                   # a plane without any weight at the OS4 scale
    ws[0][:, : H // 2] = 0                         # half of every OS1 plane empty: long borders between live and skipped tiles
    assert float(ws[0].mean()) < 0.1
    for k in range(3):
        gs = torch.rand(3, 8, generator=g).cuda()
        _check(ops, S, H, W, preds, tgt, ws, gs, f"band weights {k}")
    # isolated single-pixel weights next to tile corners (reach of the adjoint pyramid across tile boundaries)
    ws2 = [torch.zeros(S, H, W).cuda() for _ in range(2)] + [torch.ones(S, H, W).cuda()]
    for (y, x) in ((31, 31), (32, 64), (95, 33), (H - 1, W - 1), (0, 0), (64, 127)):
        ws2[0][:, y, x] = 1.0
        ws2[1][:, min(y + 1, H - 1), x] = 1.0
    _check(ops, S, H, W, preds, tgt, ws2, torch.rand(3, 8, generator=g).cuda(), "isolated weights")
