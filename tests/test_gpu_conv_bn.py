"""-m gpu: native conv+BN+act (K2 + K3 + K4, forward and backward) against the torch restatement evaluated in
fp32 on the GPU with the same fp16-rounded inputs.  Tolerances are fp16 storage noise: outputs 4e-3 of the
output scale; gradients 2e-2 of each gradient's scale (they pass through two fp16-stored tensors)."""
import pytest
import torch
from torch import nn

import ops_ref

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.float() - b.float()).abs().max()) / (float(b.float().abs().max()) + 1e-6)


def _rel_l2(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / (b.norm() + 1e-12))


CASES = [
    # N, H, W, Ci, Co, kw: k, stride, pad, dil, act, act_first, transposed, residual, res_up
    dict(N=2, H=32, W=32, Ci=64, Co=64, k=3, stride=1, pad=1, dil=1, act="relu", res=True),
    dict(N=2, H=64, W=64, Ci=32, Co=64, k=3, stride=2, pad=1, dil=1, act="relu"),
    dict(N=2, H=32, W=32, Ci=64, Co=128, k=2, stride=2, pad=0, dil=1, act=None),                 # avgpool+1x1 skip
    dict(N=1, H=64, W=64, Ci=16, Co=32, k=3, stride=1, pad=1, dil=1, act="relu", act_first=True),  # shortcut
    dict(N=2, H=16, W=16, Ci=512, Co=256, k=3, stride=1, pad=2, dil=2, act="relu"),              # ASPP dilated
    dict(N=2, H=16, W=16, Ci=256, Co=256, k=4, act="lrelu", transposed=True),                    # ConvT 4x4 s2
    dict(N=2, H=32, W=32, Ci=256, Co=128, k=3, stride=1, pad=1, dil=1, act="lrelu", res=True, res_up=True),
    dict(N=4, H=1, W=1, Ci=512, Co=256, k=1, stride=1, pad=0, dil=1, act="relu"),                # ASPP global pool
    dict(N=2, H=16, W=16, Ci=1280, Co=512, k=1, stride=1, pad=0, dil=1, act="relu"),             # ASPP fuse conv
    # the same geometries without an activation: no ReLU kink, so gradients must agree tightly
    dict(N=2, H=32, W=32, Ci=64, Co=64, k=3, stride=1, pad=1, dil=1, act=None, res=True),
    dict(N=2, H=64, W=64, Ci=32, Co=64, k=3, stride=2, pad=1, dil=1, act=None),
    dict(N=2, H=16, W=16, Ci=256, Co=256, k=4, act=None, transposed=True),
    dict(N=2, H=16, W=16, Ci=512, Co=256, k=3, stride=1, pad=4, dil=4, act=None),
    dict(N=1, H=64, W=64, Ci=16, Co=32, k=3, stride=1, pad=1, dil=1, act=None),
]


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("c", CASES, ids=lambda c: f"{c['Ci']}-{c['Co']}k{c['k']}{'T' if c.get('transposed') else ''}{c['act']}")
def test_conv_bn_act_forward_backward(c, training):
    from maggie_b200 import ops
    g = torch.Generator().manual_seed(c["Ci"] + c["Co"])
    N, H, W, Ci, Co, k = c["N"], c["H"], c["W"], c["Ci"], c["Co"], c["k"]
    tr = c.get("transposed", False)
    x = torch.randn(N, Ci, H, W, generator=g).half().cuda().contiguous(memory_format=torch.channels_last)
    wshape = (Ci, Co, k, k) if tr else (Co, Ci, k, k)
    w = (torch.randn(wshape, generator=g) / (Ci * k * k) ** 0.5).half().float().cuda()
    Ho, Wo = (2 * H, 2 * W) if tr else ((H + 2 * c["pad"] - c["dil"] * (k - 1) - 1) // c["stride"] + 1,) * 2
    if not tr:
        Wo = (W + 2 * c["pad"] - c["dil"] * (k - 1) - 1) // c["stride"] + 1
    res = None
    if c.get("res"):
        rs = (N, Co, Ho // 2, Wo // 2) if c.get("res_up") else (N, Co, Ho, Wo)
        res = torch.randn(rs, generator=g).half().cuda().contiguous(memory_format=torch.channels_last)
    kw = dict(act=c["act"], act_first=c.get("act_first", False), transposed=tr, res_up=c.get("res_up", False))
    if not tr:
        kw.update(stride=c["stride"], padding=c["pad"], dilation=c["dil"])

    def run(fn, xin, win, rin, dtype):
        bn = nn.BatchNorm2d(Co).cuda()
        with torch.no_grad():
            bn.weight.copy_(torch.linspace(0.5, 1.5, Co)), bn.bias.copy_(torch.linspace(-0.3, 0.3, Co))
            bn.running_mean.copy_(torch.linspace(-0.1, 0.1, Co)), bn.running_var.copy_(torch.linspace(0.8, 1.2, Co))
        xin = xin.to(dtype).detach().requires_grad_(True)
        win = win.detach().clone().requires_grad_(True)
        rin = rin.to(dtype).detach().requires_grad_(True) if rin is not None else None
        y = fn(xin, win, bn, training, residual=rin, **kw)
        return y, xin, win, rin, bn

    y, xg, wg, rg, bn = run(ops.conv_bn_act, x, w, res, torch.float16)
    yr, xr, wr, rr, bnr = run(ops_ref.conv_bn_act, x, w, res, torch.float32)
    assert y.shape == yr.shape and y.dtype == torch.float16
    assert _rel(y, yr) < 4e-3, f"forward rel err {_rel(y, yr)}"
    if not training:
        return
    assert _rel(bn.running_mean, bnr.running_mean) < 2e-3 and _rel(bn.running_var, bnr.running_var) < 2e-3
    assert int(bn.num_batches_tracked) == 1
    gy = torch.randn(y.shape, generator=g).half().cuda().contiguous(memory_format=torch.channels_last)
    y.backward(gy)
    yr.backward(gy.float())
    # with an activation, fp16 noise flips act' for the ~0.03 % of elements that sit on the kink (each flip is an O(1)
    # error in dz), so the bound is on the relative L2 error; without activation the agreement is tight
    tol = 4e-2 if c["act"] else 6e-3
    assert _rel_l2(xg.grad, xr.grad) < tol, f"dx rel L2 err {_rel_l2(xg.grad, xr.grad)}"
    assert _rel_l2(wg.grad, wr.grad) < tol, f"dw rel L2 err {_rel_l2(wg.grad, wr.grad)}"
    assert _rel_l2(bn.weight.grad, bnr.weight.grad) < tol and _rel_l2(bn.bias.grad, bnr.bias.grad) < tol
    if rg is not None:
        # d(residual) = dy * act'(z) element-wise: compare away from the activation kink, where fp16 noise cannot flip act'
        zr = yr.detach()
        if c.get("res_up"):
            zr = torch.nn.functional.avg_pool2d(zr.abs(), 2) * (torch.nn.functional.max_pool2d(-zr.abs(), 2) < -2e-2)
        far = zr.abs() > 2e-2
        assert float(far.float().mean()) > 0.3
        d = ((rg.grad.float() - rr.grad).abs() * far).max() / rr.grad.abs().max()
        assert float(d) < 2e-2, f"dres rel err {float(d)}"
    elif False:
        pass
