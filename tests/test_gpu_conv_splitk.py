"""EXPERIMENTAL split-K convolution (K2s, csrc/k2s_conv_splitk.cu) against torch and against K2 on the few-CTA layer shapes.
Opt-in like the kernel itself: runs only with MAGGIE_B200_CONV_SPLITK=1 (the kernel has not been validated on hardware yet
and is not part of the default path)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MAGGIE_B200_CONV_SPLITK", "0") != "1",
                                 reason="experimental kernel: set MAGGIE_B200_CONV_SPLITK=1")]


@pytest.mark.parametrize("N,H,W,Ci,Co,k,pad,dil", [
    (8, 16, 16, 512, 512, 3, 1, 1),      # bottleneck: 16 tiles x 4 N tiles, 72 k-blocks
    (8, 32, 32, 256, 256, 3, 1, 1),      # layer3: 64 tiles x 2
    (8, 16, 16, 512, 256, 3, 4, 4),      # ASPP dilated branch: 16 tiles x 2
    (8, 16, 16, 1280, 512, 1, 0, 1),     # ASPP projection, 1x1
    (2, 24, 40, 128, 128, 3, 1, 1),      # ragged grid (rows beyond the image inside a tile)
])
def test_splitk_matches_k2_and_torch(N, H, W, Ci, Co, k, pad, dil, monkeypatch):
    from maggie_b200 import dense
    g = torch.Generator().manual_seed(Ci + H)
    x = torch.randn(N, H, W, Ci, generator=g).half().cuda()
    w = (torch.randn(Co, Ci, k, k, generator=g) / (Ci * k * k) ** 0.5).half().cuda()
    bias = torch.randn(Co, generator=g).cuda()
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=pad, dilation=dil)).permute(0, 2, 3, 1)
    tol = 2e-3 * float(ref.abs().max()) + 1e-3
    outs = []
    for split in (False, True, True):            # the second split-K launch checks that the workspace was left zeroed
        monkeypatch.setattr(dense, "SPLITK", split)
        stats = dense.new_stats(Co, x.device)
        y = dense.conv2d_nhwc(x, w, padding=pad, dilation=dil, relu=True, bias=bias, stats=stats)
        assert float((y.float() - ref).abs().max()) <= tol
        s = stats.sum(0)
        n = ref.numel() / Co
        assert torch.allclose(s[0] / n, ref.mean((0, 1, 2)), atol=2e-3)
        assert torch.allclose(s[1] / n, (ref * ref).mean((0, 1, 2)), rtol=2e-3, atol=2e-3)
        outs.append(y)
    assert float((outs[0].float() - outs[1].float()).abs().max()) <= 2e-3 * float(ref.abs().max())
    assert torch.equal(outs[1], outs[2]) or float((outs[1].float() - outs[2].float()).abs().max()) <= 1e-3 * float(ref.abs().max())
    for ws in dense._SPLITK_WS.values():
        assert int(ws.count_nonzero()) == 0, "split-K workspace not left zeroed"
