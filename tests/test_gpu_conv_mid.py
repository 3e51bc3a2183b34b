"""-m gpu: K2h (csrc/k2h_conv_mid.cu) and K2t (csrc/k2t_conv_mid_t.cu, the transposed form for Co % 128 == 0), the
halo-patch / weight-streaming kernels that mg_conv_fprop routes the mid-resolution stride-1 3x3 layers to, against torch
fp32 and against the generic kernel K2 they replace."""
import pytest
import torch
import torch.nn.functional as F

from maggie_b200 import _lib, dense

pytestmark = pytest.mark.gpu


def _mid_launches():
    return _lib.lib().mg_conv_mid_launches() + _lib.lib().mg_conv_midt_launches()


def _run(x, wp, taps, hw, monkeypatch, generic, **kw):
    """generic: True = K2, False = default routing (K2t when eligible, else K2h), "h" = K2h."""
    for k in ("MAGGIE_B200_MID_CONV", "MAGGIE_B200_NO_MIDT_CONV"):
        monkeypatch.delenv(k, raising=False)
    if generic is True:
        monkeypatch.setenv("MAGGIE_B200_NO_MIDT_CONV", "1")
    elif generic == "h":
        monkeypatch.setenv("MAGGIE_B200_NO_MIDT_CONV", "1")
        monkeypatch.setenv("MAGGIE_B200_MID_CONV", "h")    # K2h is opt-in (slower than K2 / K2t since round 2)
    else:
        monkeypatch.setenv("MAGGIE_B200_MID_CONV", "h")    # default routing + K2h for what K2t does not take
    m0 = _mid_launches()
    y = dense.conv_launch(x, wp, taps, grid_hw=hw, **kw)
    return y, _mid_launches() - m0


@pytest.mark.parametrize("shape", [
    # N, H, W, Ci, Co
    (8, 64, 64, 128, 128), (8, 32, 32, 256, 256), (8, 16, 16, 512, 512), (2, 64, 64, 256, 128), (3, 32, 32, 512, 256),
    (1, 24, 40, 128, 192), (2, 60, 52, 128, 64), (1, 8, 8, 128, 128), (5, 13, 17, 192, 128),
])
@pytest.mark.parametrize("variant", [False, "h"])
def test_mid_conv_matches_torch_and_generic(shape, variant, monkeypatch):
    N, H, W, Ci, Co = shape
    g = torch.Generator(device="cuda").manual_seed(N * H + Ci)
    x = torch.randn(N, H, W, Ci, device="cuda", generator=g).half()
    w = torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / (Ci * 9) ** 0.5
    wp, taps = dense.pack_weight(w, Ci), dense.conv_taps(3, 3, 1, 1, Ci)
    t0 = _lib.lib().mg_conv_midt_launches()
    y, used = _run(x, wp, taps, (H, W), monkeypatch, variant)
    assert used == 1, "layer was not routed to K2h / K2t"
    took_t = _lib.lib().mg_conv_midt_launches() - t0
    if variant == "h" or Co % 128:
        assert took_t == 0
    elif shape in ((8, 64, 64, 128, 128), (8, 32, 32, 256, 256), (8, 16, 16, 512, 512)):
        assert took_t == 1, "K2t did not take a layer it is built for"
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
    for generic in (False, "h", True):
        st = torch.zeros(dense.STAT_COPIES, 2, Co, device="cuda")
        y, used = _run(x, wp, taps, (H, W), monkeypatch, generic, stats=st, pre_act="relu")
        assert used == (0 if generic is True else 1)
        out.append((y.float(), st.sum(0)))
    (y1, s1), (y2, s2), (y0, s0) = out
    for yy, ss in ((y1, s1), (y2, s2)):
        assert float((yy - y0).abs().max()) < 1e-3 * float(y0.abs().max())
        assert float((ss - s0).abs().max()) < 1e-4 * float(s0.abs().max())
    # eval forward: bias, affine, residual, LeakyReLU after the residual
    ys = [_run(x, wp, taps, (H, W), monkeypatch, generic, bias=bias, scale=sc, shift=sh, res=res, post_act="lrelu")[0].float()
          for generic in (False, "h", True)]
    assert float((ys[0] - ys[2]).abs().max()) < 2e-3 * float(ys[2].abs().max())
    assert float((ys[1] - ys[2]).abs().max()) < 2e-3 * float(ys[2].abs().max())


def test_midt_conv_concat_slice_and_two_items_per_cta(monkeypatch):
    """K2t: output into a channel slice of a wider buffer (TMA store at a channel offset), more items than SMs (the
    persistent loop re-uses the patch buffer as output staging), an image height the slab height does not divide."""
    g = torch.Generator(device="cuda").manual_seed(5)
    N, H, W, Ci, Co = 20, 30, 32, 128, 256
    x = torch.randn(N, H, W, Ci, device="cuda", generator=g).half()
    w = torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / (Ci * 9) ** 0.5
    wp, taps = dense.pack_weight(w, Ci), dense.conv_taps(3, 3, 1, 1, Ci)
    buf = torch.full((N, H, W, Co + 64), 7.0, device="cuda", dtype=torch.float16)
    t0 = _lib.lib().mg_conv_midt_launches()
    _run(x, wp, taps, (H, W), monkeypatch, False, out=buf, c_off=32)
    assert _lib.lib().mg_conv_midt_launches() == t0 + 1
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.half().float(), padding=1).permute(0, 2, 3, 1)
    assert float((buf[..., 32:32 + Co].float() - ref).abs().max()) < 2e-3 * float(ref.abs().max())
    assert bool((buf[..., :32] == 7).all()) and bool((buf[..., 32 + Co:] == 7).all())


def test_mid_conv_serves_the_data_gradient(monkeypatch):
    """dgrad of a 3x3 stride-1 layer = the same kernel with the transposed pack and mirrored taps."""
    g = torch.Generator(device="cuda").manual_seed(3)
    N, H, W, Ci, Co = 2, 32, 32, 128, 256
    dy = torch.randn(N, H, W, Co, device="cuda", generator=g).half()
    w = torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / (Ci * 9) ** 0.5
    geom = dense.ConvGeom("conv", 3, 1, 1, 1)
    monkeypatch.delenv("MAGGIE_B200_NO_MIDT_CONV", raising=False)
    m0 = _mid_launches()
    dx = geom.dgrad(dy, w, (N, H, W, Ci))
    assert _mid_launches() == m0 + 1
    ref = F.conv_transpose2d(dy.permute(0, 3, 1, 2).float(), w.half().float(), padding=1).permute(0, 2, 3, 1)
    assert float((dx.float() - ref).abs().max()) < 2e-3 * float(ref.abs().max())
