"""-m gpu: the tcgen05/TMA implicit-GEMM convolution (K2) through the C ABI against torch's conv evaluated in
fp32 on the same fp16-valued inputs.  Tolerance: fp16 rounding of the stored output (2^-11 relative) plus
fp32 accumulation-order noise: |diff| <= 2e-3 * max|ref| + 1e-3."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close(got, ref, what=""):
    tol = 2e-3 * float(ref.abs().max()) + 1e-3
    err = float((got.float() - ref).abs().max())
    assert err <= tol, f"{what}: max err {err:.4e} > tol {tol:.4e}"


def _mk(N, H, W, Ci, Co, k, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(N, H, W, Ci, generator=g)).half().cuda()
    w = (torch.randn(Co, Ci, k, k, generator=g) / (Ci * k * k) ** 0.5).half().cuda()
    return x, w


@pytest.mark.parametrize("N,H,W,Ci,Co,k,stride,pad,dil", [
    (2, 32, 32, 64, 64, 3, 1, 1, 1),      # 128B swizzle, one K chunk per tap
    (1, 64, 64, 16, 32, 3, 1, 1, 1),      # 32B swizzle (first layer: 6 -> 16 padded channels)
    (2, 64, 64, 32, 64, 3, 2, 1, 1),      # 64B swizzle, stride 2 via TMA elementStrides
    (2, 16, 16, 128, 256, 1, 1, 0, 1),    # 1x1, two N tiles, two K chunks
    (1, 16, 16, 512, 256, 3, 1, 4, 4),    # dilated (ASPP)
    (1, 24, 40, 64, 128, 3, 1, 1, 1),     # grid not a multiple of the 8x16 tile
    (3, 8, 8, 256, 512, 3, 1, 1, 1),      # narrow image -> 16x8 tile, four N tiles
    (2, 32, 32, 256, 512, 3, 2, 1, 1),    # stride 2, deep K
    (8, 1, 1, 512, 256, 1, 1, 0, 1),      # ASPP global-pool branch: 1x1 spatial
    (1, 128, 128, 32, 32, 3, 1, 1, 1),
])
def test_conv_fprop_matches_torch(N, H, W, Ci, Co, k, stride, pad, dil):
    from maggie_b200 import dense
    x, w = _mk(N, H, W, Ci, Co, k, seed=H + Ci)
    got = dense.conv2d_nhwc(x, w, stride=stride, padding=pad, dilation=dil)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), stride=stride, padding=pad, dilation=dil).permute(0, 2, 3, 1)
    assert got.shape == ref.shape
    _close(got, ref, f"conv {Ci}->{Co} k{k} s{stride} d{dil}")


def test_conv_relu_bias_stats_and_channel_slice():
    from maggie_b200 import dense
    x, w = _mk(2, 40, 24, 64, 64, 3, seed=5)
    bias = torch.randn(64).cuda()
    stats = dense.new_stats(64, x.device)
    out = torch.zeros(2, 40, 24, 192, dtype=torch.float16, device="cuda")
    dense.conv2d_nhwc(x, w, relu=True, bias=bias, stats=stats, out=out, c_off=64)
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1)).permute(0, 2, 3, 1)
    _close(out[..., 64:128], ref, "relu+bias into channel slice")
    assert float(out[..., :64].abs().max()) == 0 and float(out[..., 128:].abs().max()) == 0
    s = stats.sum(0)
    n = ref.numel() / 64
    assert torch.allclose(s[0] / n, ref.mean((0, 1, 2)), atol=2e-3)
    assert torch.allclose(s[1] / n, (ref * ref).mean((0, 1, 2)), rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("N,H,W,Ci,Co", [(2, 16, 16, 512, 512), (1, 32, 32, 256, 256), (2, 8, 8, 64, 64)])
def test_conv_transpose_4x4_s2(N, H, W, Ci, Co):
    from maggie_b200 import dense
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, H, W, Ci, generator=g).half().cuda()
    w = (torch.randn(Ci, Co, 4, 4, generator=g) / (Ci * 4) ** 0.5).half().cuda()
    got = dense.conv_transpose4x4s2_nhwc(x, w)
    ref = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), w.float(), stride=2, padding=1).permute(0, 2, 3, 1)
    assert got.shape == ref.shape
    _close(got, ref, "convT 4x4 s2")


def test_conv_full_size_linearity():
    """C2-size layer (8 x 256 x 256 x 32 -> 32): conv(a*x1 + x2) == a*conv(x1) + conv(x2) up to fp16 rounding, and
    agreement with torch on a crop."""
    from maggie_b200 import dense
    x1, w = _mk(8, 256, 256, 32, 32, 3, seed=11)
    x2 = torch.randn_like(x1)
    y1, y2 = dense.conv2d_nhwc(x1, w).float(), dense.conv2d_nhwc(x2, w).float()
    y12 = dense.conv2d_nhwc((2 * x1 + x2), w).float()
    assert float((y12 - (2 * y1 + y2)).abs().max()) < 2e-2
    ref = F.conv2d(x1[3:4, :64].float().permute(0, 3, 1, 2), w.float(), padding=1).permute(0, 2, 3, 1)
    _close(y1[3:4, :62], ref[:, :62], "crop")


# ---- K2b: halo-resident persistent kernel for the high-resolution, low-channel stride-1 layers --------------------
@pytest.mark.parametrize("N,H,W,Ci,Co", [
    (2, 40, 72, 32, 32),       # one strip (W <= 254), ragged last row block
    (1, 128, 128, 64, 64),     # 128B-swizzled weights, 8 channel planes
    (1, 24, 300, 32, 64),      # three 128-pixel strips, the last one partial
    (8, 64, 64, 32, 64),
    (1, 9, 64, 64, 32),        # fewer rows than a row block
    (2, 512, 512, 32, 32),     # the heaviest C2 layer
])
def test_halo_conv_matches_torch_and_generic_kernel(N, H, W, Ci, Co, monkeypatch):
    from maggie_b200 import _lib, dense
    x, w = _mk(N, H, W, Ci, Co, 3, seed=H + W + Ci)
    before = _lib.lib().mg_conv_halo_launches()
    got = dense.conv2d_nhwc(x, w)
    assert _lib.lib().mg_conv_halo_launches() == before + 1, "the layer was not routed to the halo kernel"
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), padding=1).permute(0, 2, 3, 1)
    _close(got, ref, f"halo conv {Ci}->{Co} {H}x{W}")
    monkeypatch.setenv("MAGGIE_B200_NO_HALO_CONV", "1")
    gen = dense.conv2d_nhwc(x, w)
    assert _lib.lib().mg_conv_halo_launches() == before + 1
    # same products, fp32 accumulation in a different order: equal up to one fp16 ulp of the output
    assert float((got.float() - gen.float()).abs().max()) <= 2e-3 * float(ref.abs().max()) + 1e-3


def test_halo_conv_epilogue_stats_bias_act_residual():
    from maggie_b200 import _lib, dense
    x, w = _mk(2, 48, 136, 32, 32, 3, seed=11)
    bias = torch.randn(32).cuda()
    stats = dense.new_stats(32, x.device)
    before = _lib.lib().mg_conv_halo_launches()
    out = dense.conv2d_nhwc(x, w, bias=bias, stats=stats, pre_act="lrelu")
    assert _lib.lib().mg_conv_halo_launches() == before + 1
    ref = F.leaky_relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1), 0.2).permute(0, 2, 3, 1)
    _close(out, ref, "bias + leaky relu")
    s, n = stats.sum(0), ref.numel() / 32
    assert torch.allclose(s[0] / n, ref.mean((0, 1, 2)), atol=2e-3)
    assert torch.allclose(s[1] / n, (ref * ref).mean((0, 1, 2)), rtol=2e-3, atol=2e-3)
    # eval-mode epilogue: affine + residual + ReLU
    scale, shift = torch.rand(32).cuda() + 0.5, torch.randn(32).cuda()
    res = torch.randn(2, 48, 136, 32).half().cuda()
    out2 = dense.conv2d_nhwc(x, w, scale=scale, shift=shift, res=res, post_act="relu")
    conv = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), padding=1).permute(0, 2, 3, 1)
    _close(out2, F.relu(conv * scale + shift + res.float()), "affine + residual + relu")


def test_halo_conv_data_gradient_path():
    """The stride-1 dgrad (flipped taps, transposed pack) of a 32 -> 32 layer also runs on the halo kernel."""
    from maggie_b200 import _lib, dense
    x, w = _mk(2, 64, 128, 32, 32, 3, seed=3)
    dy = torch.randn(2, 64, 128, 32).half().cuda()
    g = dense.ConvGeom("conv", 3, 1, 1, 1)
    before = _lib.lib().mg_conv_halo_launches()
    dx = g.dgrad(dy, w.float(), x.shape)
    assert _lib.lib().mg_conv_halo_launches() == before + 1
    ref = torch.nn.grad.conv2d_input((2, 32, 64, 128), w.float(), dy.float().permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
    _close(dx, ref, "halo dgrad")


# ---- K4b: halo-resident persistent weight gradient of the high-resolution 32- / 64-channel 3x3 layers --------------
@pytest.mark.parametrize("N,H,W,Co,Ci", [(2, 256, 256, 32, 32), (1, 512, 256, 64, 32), (3, 130, 128, 32, 32), (2, 256, 384, 16, 32),
                                         (8, 128, 128, 64, 64), (2, 256, 128, 32, 64), (3, 130, 128, 48, 64)])
def test_halo_wgrad_matches_torch_and_generic_kernel(N, H, W, Co, Ci, monkeypatch):
    """Ci = 64: two CTA populations, each accumulating 32 of the input channels (the [128 x 9*64] tile exceeds TMEM)."""
    from maggie_b200 import _lib, dense
    g = torch.Generator().manual_seed(H + W + Co + Ci)
    x = torch.randn(N, H, W, Ci, generator=g).half().cuda()
    dy = (torch.randn(N, H, W, Co, generator=g) * 0.1).half().cuda()
    taps = dense.conv_taps(3, 3, 1, 1, Ci)
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Co, Ci, 3, 3), dy.float().permute(0, 3, 1, 2), padding=1)
    outs = []
    for mode in ("halo", "halo_one_tap_per_mma", "generic"):
        monkeypatch.setenv("MAGGIE_B200_NO_HALO_CONV", "1" if mode == "generic" else "0")
        monkeypatch.setenv("MAGGIE_B200_WGRAD_HALO_MODE", "1" if mode == "halo_one_tap_per_mma" else "0")
        dw = torch.zeros(Co, 9 * Ci, device="cuda")
        before = _lib.lib().mg_wgrad_halo_launches()
        dense.wgrad_launch(dy, x, taps, dw, grid_hw=(H, W))
        # (persistent kernel: taken when there are at least two work items per SM)
        routed = mode != "generic" and N * ((H + 1) // 2) * (W // 128) * (Ci // 32) >= 2 * 148
        assert _lib.lib().mg_wgrad_halo_launches() == before + routed, "routing"
        got = dw.view(Co, 3, 3, Ci).permute(0, 3, 1, 2)
        assert float((got - ref).abs().max()) <= 2e-3 * float(ref.abs().max()) + 1e-3, mode
        outs.append(got)
    # same products, different summation order (split-K partition): tiny differences only
    assert float((outs[0] - outs[2]).abs().max()) <= 1e-3 * float(ref.abs().max()) + 1e-3


# ---- K4: split-K flush through TMA reduce-add (rows beyond Co clipped, several channel tiles / tap groups) ---------------
@pytest.mark.parametrize("N,H,W,Ci,Co", [(8, 64, 64, 128, 128), (2, 32, 32, 256, 256), (1, 16, 16, 512, 512), (2, 128, 128, 64, 64),
                                         (1, 24, 40, 128, 192), (3, 13, 17, 128, 64), (2, 64, 64, 256, 128), (1, 20, 52, 64, 32),
                                         (1, 40, 24, 16, 48)])
def test_wgrad_reduce_flush_matches_torch(N, H, W, Ci, Co):
    from maggie_b200 import dense
    g = torch.Generator().manual_seed(H + W + Co + Ci)
    x = torch.randn(N, H, W, Ci, generator=g).half().cuda()
    dy = (torch.randn(N, H, W, Co, generator=g) * 0.1).half().cuda()
    taps = dense.conv_taps(3, 3, 1, 1, Ci)
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Co, Ci, 3, 3), dy.float().permute(0, 3, 1, 2), padding=1)
    dw = torch.full((Co, 9 * Ci), 0.5, device="cuda")                  # accumulates INTO the buffer
    dense.wgrad_launch(dy, x, taps, dw, grid_hw=(H, W))
    got = (dw - 0.5).view(Co, 3, 3, Ci).permute(0, 3, 1, 2)
    assert float((got - ref).abs().max()) <= 2e-3 * float(ref.abs().max()) + 1e-3
