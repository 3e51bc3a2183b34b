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
