"""-m gpu: K2h (csrc/k2h_conv_mid.cu), the halo-patch / weight-streaming kernel that mg_conv_fprop routes the mid-resolution
stride-1 3x3 layers to, against torch fp32 and against the generic kernel K2 it replaces."""
import pytest
import torch
import torch.nn.functional as F

from maggie_b200 import _lib, dense

pytestmark = pytest.mark.gpu


def _run(x, wp, taps, hw, monkeypatch, generic, **kw):
    if generic:
        monkeypatch.setenv("MAGGIE_B200_NO_MID_CONV", "1")
    else:
        monkeypatch.delenv("MAGGIE_B200_NO_MID_CONV", raising=False)
    m0 = _lib.lib().mg_conv_mid_launches()
    y = dense.conv_launch(x, wp, taps, grid_hw=hw, **kw)
    return y, _lib.lib().mg_conv_mid_launches() - m0


@pytest.mark.parametrize("shape", [
    # N, H, W, Ci, Co
    (8, 64, 64, 128, 128), (8, 32, 32, 256, 256), (8, 16, 16, 512, 512), (2, 64, 64, 256, 128), (3, 32, 32, 512, 256),
    (1, 24, 40, 128, 192), (2, 60, 52, 128, 64), (1, 8, 8, 128, 128), (5, 13, 17, 192, 128),
])
def test_mid_conv_matches_torch_and_generic(shape, monkeypatch):
    N, H, W, Ci, Co = shape
    g = torch.Generator(device="cuda").manual_seed(N * H + Ci)
    x = torch.randn(N, H, W, Ci, device="cuda", generator=g).half()
    w = torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / (Ci * 9) ** 0.5
    wp, taps = dense.pack_weight(w, Ci), dense.conv_taps(3, 3, 1, 1, Ci)
    y, used = _run(x, wp, taps, (H, W), monkeypatch, False)
    assert used == 1, "layer was not routed to K2h"
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.half().float(), padding=1).permute(0, 2, 3, 1)
    scale = float(ref.abs().max())
    assert float((y.float() - ref).abs().max()) < 2e-3 * scale      # fp16 output rounding
    if Co % 128 == 0 or Co < 128:                                   # (the generic kernel needs Co to fill its N tiles)
        y0, used0 = _run(x, wp, taps, (H, W), monkeypatch, True)
        assert used0 == 0
        assert float((y.float() - y0.float()).abs().max()) < 1e-3 * scale  # same products, different fp32 summation order


def test_mid_conv_epilogues_match_generic(monkeypatch):
    g = torch.Generator(device="cuda").manual_seed(77)
    N, H, W, Ci, Co = 4, 32, 32, 256, 128
    x = torch.randn(N, H, W, Ci, device="cuda", generator=g).half()
    w = torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / (Ci * 9) ** 0.5
    res = torch.randn(N, H, W, Co, device="cuda", generator=g).half()
    bias, sc, sh = (torch.randn(Co, device="cuda", generator=g) for _ in range(3))
    wp, taps = dense.pack_weight(w, Ci), dense.conv_taps(3, 3, 1, 1, Ci)
    # training forward: ReLU before the statistics (shortcut branches), statistics of the fp32 pre-images
    out = []
    for generic in (False, True):
        st = torch.zeros(dense.STAT_COPIES, 2, Co, device="cuda")
        y, used = _run(x, wp, taps, (H, W), monkeypatch, generic, stats=st, pre_act="relu")
        assert used == (0 if generic else 1)
        out.append((y.float(), st.sum(0)))
    (y1, s1), (y0, s0) = out
    assert float((y1 - y0).abs().max()) < 1e-3 * float(y0.abs().max())
    assert float((s1 - s0).abs().max()) < 1e-4 * float(s0.abs().max())
    # eval forward: bias, affine, residual, LeakyReLU after the residual
    ys = [_run(x, wp, taps, (H, W), monkeypatch, generic, bias=bias, scale=sc, shift=sh, res=res, post_act="lrelu")[0].float()
          for generic in (False, True)]
    assert float((ys[0] - ys[1]).abs().max()) < 2e-3 * float(ys[1].abs().max())


def test_mid_conv_serves_the_data_gradient(monkeypatch):
    """dgrad of a 3x3 stride-1 layer = the same kernel with the transposed pack and mirrored taps."""
    g = torch.Generator(device="cuda").manual_seed(3)
    N, H, W, Ci, Co = 2, 32, 32, 128, 256
    dy = torch.randn(N, H, W, Co, device="cuda", generator=g).half()
    w = torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / (Ci * 9) ** 0.5
    geom = dense.ConvGeom("conv", 3, 1, 1, 1)
    monkeypatch.delenv("MAGGIE_B200_NO_MID_CONV", raising=False)
    m0 = _lib.lib().mg_conv_mid_launches()
    dx = geom.dgrad(dy, w, (N, H, W, Ci))
    assert _lib.lib().mg_conv_mid_launches() == m0 + 1
    ref = F.conv_transpose2d(dy.permute(0, 3, 1, 2).float(), w.half().float(), padding=1).permute(0, 2, 3, 1)
    assert float((dx.float() - ref).abs().max()) < 2e-3 * float(ref.abs().max())
