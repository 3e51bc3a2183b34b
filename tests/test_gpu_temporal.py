"""-m gpu: K11 (csrc/k11_temporal.cu) - the ConvGRU step with its gate kernels around the native convolutions and the
bidirectional temporal fusion - against the torch restatement of module/conv_gru.py:50-58 and
decoder/resnet_inst_matt_spconv_temp.py:122-142 (fp32), forward and every gradient."""
import pytest
import torch
import torch.nn.functional as F

from maggie_b200 import _lib, ops

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))


def _gru_ref(x, h, w_ih, b_ih, w_hh, b_hh):
    C = x.shape[1]
    rz = torch.sigmoid(F.conv2d(torch.cat([x, h], 1), w_ih, b_ih, padding=1))
    r, z = rz[:, :C], rz[:, C:]
    c = torch.tanh(F.conv2d(torch.cat([x, r * h], 1), w_hh, b_hh, padding=1))
    return (1 - z) * h + z * c


@pytest.mark.parametrize("N,C,hh,ww", [(2, 128, 60, 104), (3, 32, 16, 24)])
def test_gru_step_fwd_bwd(N, C, hh, ww):
    g = torch.Generator().manual_seed(C + hh)
    x = (torch.randn(N, C, hh, ww, generator=g) * 0.7).cuda()
    h = (torch.randn(N, C, hh, ww, generator=g) * 0.7).cuda()
    w_ih = (torch.randn(2 * C, 2 * C, 3, 3, generator=g) / (18 * C) ** 0.5).cuda()
    w_hh = (torch.randn(C, 2 * C, 3, 3, generator=g) / (18 * C) ** 0.5).cuda()
    b_ih, b_hh = (torch.randn(2 * C, generator=g) * 0.2).cuda(), (torch.randn(C, generator=g) * 0.2).cuda()
    gy = torch.randn(N, C, hh, ww, generator=g).cuda()
    cl = lambda t: t.half().contiguous(memory_format=torch.channels_last)
    ours = [t.detach().clone().requires_grad_(True) for t in (cl(x), cl(h), w_ih, b_ih, w_hh, b_hh)]
    n0 = _lib.launch_count()
    y = ops.gru_step(*ours)
    assert y.dtype == torch.float16 and _lib.launch_count() - n0 == 5      # concat, conv, gate1, conv, gate2: no torch glue
    y.backward(cl(gy))
    # reference on the same fp16-rounded operands, fp32 arithmetic
    ref = [t.detach().clone().requires_grad_(True) for t in (cl(x).float(), cl(h).float(), w_ih.half().float(), b_ih, w_hh.half().float(), b_hh)]
    yr = _gru_ref(*ref)
    yr.backward(cl(gy).float())
    assert _rel(y, yr) < 3e-3
    for a, b, name in zip(ours, ref, ("dx", "dh", "dw_ih", "db_ih", "dw_hh", "db_hh")):
        assert _rel(a.grad, b.grad) < 1e-2, (name, _rel(a.grad, b.grad))


def _fuse_ref(fd, bd, preds):
    n_f = preds.shape[1]
    fp = [preds[:, 0]]
    for i in range(1, n_f):
        s = torch.sigmoid(fd[:, i])
        fp.append(fp[-1] * (1 - s) + preds[:, i] * s)
    bp = [preds[:, n_f - 1]]
    for i in range(n_f - 1, 0, -1):
        s = torch.sigmoid(bd[:, i - 1])
        bp.append(bp[-1] * (1 - s) + preds[:, i - 1] * s)
    bp = bp[::-1]
    return torch.stack([fp[0]] + [(fp[i] + bp[i]) / 2 for i in range(1, n_f - 1)] + [bp[n_f - 1]], 1)


@pytest.mark.parametrize("B,n_f,n_i,H,W", [(1, 5, 2, 96, 160), (2, 3, 3, 40, 56), (1, 2, 1, 16, 24), (1, 8, 10, 24, 32)])
def test_temporal_fuse_fwd_bwd(B, n_f, n_i, H, W):
    g = torch.Generator().manual_seed(n_f * 7 + n_i)
    fd = (torch.randn(B, n_f, 1, H, W, generator=g) * 2).cuda()
    bd = (torch.randn(B, n_f, 1, H, W, generator=g) * 2).cuda()
    preds = torch.rand(B, n_f, n_i, H, W, generator=g).cuda()
    gy = torch.randn(B, n_f, n_i, H, W, generator=g).cuda()
    a = [t.detach().clone().requires_grad_(True) for t in (fd, bd, preds)]
    b = [t.detach().clone().double().requires_grad_(True) for t in (fd, bd, preds)]
    n0 = _lib.launch_count()
    ya = ops.temporal_fuse(*a)
    assert _lib.launch_count() - n0 == 1
    yb = _fuse_ref(*b)
    assert float((ya.double() - yb).abs().max()) < 2e-6
    ya.backward(gy)
    yb.backward(gy.double())
    for x, y, name in zip(a, b, ("dfd", "dbd", "dpreds")):
        gref = y.grad.clone() if y.grad is not None else torch.zeros_like(y)   # (n_f = 2: the output is [p_0, p_1])
        if name == "dfd":
            gref[:, 0] = 0         # the reference's zero planes: no gradient defined through them
        if name == "dbd":
            gref[:, -1] = 0
        assert float((x.grad.double() - gref).abs().max()) < 1e-5 * max(1.0, float(gref.abs().max())), name
