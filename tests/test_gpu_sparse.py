"""-m gpu: K9 sparse kernels (gather -> tcgen05 tile -> rows, wgrad, heads, dense<->sparse gathers) through the C
ABI against the torch restatements in tests/ops_ref.py, on real rulebook tables from K8b.
Tolerances: fp16 storage noise (outputs 4e-3 of scale; gradients rel-L2 1e-2 without activation, 4e-2 with)."""
import numpy as np
import pytest
import torch
from torch import nn

import ops_ref
from oracle import synth
from oracle import unknown as U

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).abs().max() / (b.abs().max() + 1e-6))


def _rel_l2(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.fixture(scope="module")
def tables():
    from maggie_b200 import ops
    al = synth.soft_ellipse_alphas(2, 3, 128, 192, edge_px=4.0).numpy().reshape(6, 128, 192)
    roi = torch.from_numpy(U.compute_unknown(al, [15] * 6)).cuda()
    T = ops.build_sites(roi)
    assert T.counts[0] > 3000 and T.counts[3] > 50
    return T


def _w(co, k, ci, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(co, k, k, ci, generator=g) / (ci * k * k) ** 0.5).half().float().cuda()


CASES = [
    # name, level of the output sites, Cin, Cout, k, kind, mode, act, bias
    ("subm64", 2, 64, 64, 3, "subm", "plain", None, False),
    ("subm64_bias", 2, 64, 64, 3, "subm", "plain", None, True),
    ("subm32", 0, 32, 32, 3, "subm", "plain", None, False),
    ("subm64to32_bn", 2, 64, 32, 3, "subm", "bn_act", "lrelu", False),
    ("inv64_bn", 2, 64, 64, 3, "inv", "bn_act", "lrelu", False),
    ("inv64to32_bn", 1, 64, 32, 3, "inv", "bn_act", "lrelu", False),
    ("inv32_bn", 0, 32, 32, 3, "inv", "bn_act", "lrelu", False),
    ("pw128to64_bn", 2, 128, 64, 1, "pw", "bn_act", "lrelu", False),
    ("pw64to32_actbn", 1, 64, 32, 1, "pw", "act_bn", "relu", True),
    ("pw32", 1, 32, 32, 1, "pw", "plain", None, False),
    ("inv64_plain", 2, 64, 64, 3, "inv", "plain", None, False),
]


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_rows_conv_forward_backward(tables, case, training):
    from maggie_b200 import ops
    name, lvl, ci, co, k, kind, mode, act, has_bias = case
    T = tables
    if kind == "subm":
        table = table_t = T.nbr[lvl]
        n_src, mirror = T.counts[lvl], True
    elif kind == "inv":
        table, table_t = T.parent[lvl], T.child[lvl + 1]
        n_src, mirror = T.counts[lvl + 1], False
    else:
        table = table_t = None
        n_src, mirror = T.counts[lvl], False
    g = torch.Generator().manual_seed(ci + co + k)
    src = torch.randn(n_src, ci, generator=g).half().cuda()
    w = _w(co, k, ci, ci * co)
    bias = (torch.randn(co, generator=g) * 0.3).cuda() if has_bias else None

    def run(fn, dtype):
        bn = None
        if mode != "plain":
            bn = nn.BatchNorm1d(co).cuda()
            with torch.no_grad():
                bn.weight.copy_(torch.linspace(0.5, 1.5, co)), bn.bias.copy_(torch.linspace(-0.3, 0.3, co))
                bn.running_mean.copy_(torch.linspace(-0.1, 0.1, co)), bn.running_var.copy_(torch.linspace(0.8, 1.2, co))
        s = src.to(dtype).detach().requires_grad_(True)
        ww = w.detach().clone().requires_grad_(True)
        bb = bias.detach().clone().requires_grad_(True) if bias is not None else None
        y = fn(s, ww, bb, table=table, table_t=table_t, mirror=mirror, bn=bn, mode=mode, act=act, training=training)
        return y, s, ww, bb, bn

    y, s, ww, bb, bn = run(ops.rows_conv, torch.float16)
    yr, sr, wr, br, bnr = run(ops_ref.rows_conv, torch.float32)
    assert y.shape == yr.shape == (T.counts[lvl], co) and y.dtype == torch.float16
    assert _rel(y, yr) < 4e-3, f"forward rel err {_rel(y, yr)}"
    if not training or (mode != "plain" and False):
        return
    gy = torch.randn(y.shape, generator=g).half().cuda()
    y.backward(gy)
    yr.backward(gy.float())
    tol = 4e-2 if act else 1e-2
    assert _rel_l2(s.grad, sr.grad) < tol, f"dsrc {_rel_l2(s.grad, sr.grad)}"
    assert _rel_l2(ww.grad, wr.grad) < tol, f"dw {_rel_l2(ww.grad, wr.grad)}"
    if bb is not None:
        assert _rel_l2(bb.grad, br.grad) < tol
    if bn is not None:
        assert _rel_l2(bn.weight.grad, bnr.weight.grad) < tol and _rel_l2(bn.bias.grad, bnr.bias.grad) < tol
        assert _rel(bn.running_var, bnr.running_var) < 3e-3


@pytest.mark.parametrize("lvl,ci", [(2, 32), (0, 32)])
def test_rows_head_logit_map(tables, lvl, ci):
    from maggie_b200 import ops
    T = tables
    H, W = T.shapes[lvl]
    g = torch.Generator().manual_seed(lvl)
    src = torch.randn(T.counts[lvl], ci, generator=g).half().cuda()
    w = _w(1, 3, ci, 7)
    bias = torch.tensor([0.25]).cuda()

    def run(fn, dtype):
        s = src.to(dtype).detach().requires_grad_(True)
        ww, bb = w.detach().clone().requires_grad_(True), bias.detach().clone().requires_grad_(True)
        return fn(s, ww, bb, T.nbr[lvl], T.coords[lvl], 6, H, W), s, ww, bb

    m, s, ww, bb = run(ops.rows_head, torch.float16)
    mr, sr, wr, br = run(ops_ref.rows_head, torch.float32)
    assert m.shape == (6, 1, H, W) and m.dtype == torch.float32
    assert bool(((m == -99.0) == (mr == -99.0)).all())
    assert float((m - mr).abs().max()) < 5e-3 * float(mr[mr != -99].abs().max()) + 1e-3
    gm = torch.randn(m.shape, generator=g).cuda()
    m.backward(gm)
    mr.backward(gm)
    assert _rel_l2(s.grad, sr.grad) < 1e-2 and _rel_l2(ww.grad, wr.grad) < 1e-2 and _rel_l2(bb.grad, br.grad) < 1e-2


def test_gather_dense_and_scatter_back(tables):
    from maggie_b200 import ops
    T = tables
    n_i = 3
    H, W = T.shapes[1]
    dense = torch.randn(2, 32, H, W).half().cuda().contiguous(memory_format=torch.channels_last)
    d1 = dense.detach().clone().requires_grad_(True)
    d2 = dense.detach().float().requires_grad_(True)
    a = ops.gather_dense(d1, T.coords[1], n_i)
    b = ops_ref.gather_dense(d2, T.coords[1], n_i)
    assert bool((a.float() == b).all())      # a gather is exact
    gy = torch.randn(a.shape).half().cuda()
    a.backward(gy)
    b.backward(gy.float())
    assert _rel(d1.grad, d2.grad) < 2e-3     # up to n_i fp16 additions per pixel


@pytest.mark.parametrize("ci,co", [(64, 32), (32, 64), (64, 64)])
def test_pointwise_rows_on_large_site_lists_persistent_vs_generic(ci, co, monkeypatch):
    """1x1 / Linear layers on a large site list (layer5_smooth / layer4_smooth shapes): 64 input channels run on the
    persistent kernels as two virtual taps of 32 channels; forward and both gradients must match the generic kernels."""
    from maggie_b200 import ops
    g = torch.Generator().manual_seed(ci + co)
    N = 60000                                           # >= 2 * 148 tiles of 128 rows
    src = (torch.randn(N, ci, generator=g) * 0.5).half().cuda()
    w = (torch.randn(co, 1, 1, ci, generator=g) / ci ** 0.5).cuda()
    b = torch.randn(co, generator=g).cuda()
    gy = torch.randn(N, co, generator=g).half().cuda()
    res = []
    for generic in (False, True):
        if generic:
            monkeypatch.setenv("MAGGIE_B200_NO_PERSISTENT_SPARSE", "1")
        else:
            monkeypatch.delenv("MAGGIE_B200_NO_PERSISTENT_SPARSE", raising=False)
        s, ww, bb = src.detach().clone().requires_grad_(True), w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
        y = ops.rows_conv(s, ww, bb)
        y.backward(gy)
        res.append((y.detach().float(), s.grad.float(), ww.grad.float(), bb.grad.float()))
    ref = src.float() @ w.reshape(co, ci).t() + b
    assert _rel_l2(res[0][0], ref) < 2e-3
    for a, c, what in zip(res[0], res[1], ("y", "dsrc", "dw", "dbias")):
        assert _rel_l2(a, c) < 2e-3, f"{what}: persistent vs generic {_rel_l2(a, c)}"


@pytest.mark.parametrize("mode,training", [("plain", False), ("bn_act", True)])
def test_dense_rows_take_the_conv_kernels(mode, training, monkeypatch):
    """Linear layers over contiguous dense rows (the pixel-side attention projections, 128 channels on 8 x 64 x 64 rows) are
    routed to the 1x1 case of the dense conv / weight-gradient kernels: same results as the rulebook kernels."""
    from maggie_b200 import _lib, ops, sparse
    g = torch.Generator().manual_seed(11)
    N, ci, co = 32768, 128, 128
    src = (torch.randn(N, ci, generator=g) * 0.5).half().cuda()
    w = (torch.randn(co, ci, generator=g) / ci ** 0.5).cuda()
    b = torch.randn(co, generator=g).cuda()
    gy = torch.randn(N, co, generator=g).half().cuda()
    res = []
    for dense_route in (True, False):
        monkeypatch.setattr(sparse, "DENSE_ROWS", dense_route)
        bn = torch.nn.BatchNorm1d(co).cuda() if mode != "plain" else None
        s, ww, bb = (t.detach().clone().requires_grad_(True) for t in (src, w, b))
        n0 = _lib.launch_count()
        if mode == "plain":
            y = ops.rows_conv(s, ww, bb)
        else:
            y = ops.rows_conv(s, ww, None, bn=bn, mode=mode, act="relu", training=training)
        y.backward(gy)
        res.append([y.detach().float(), s.grad.float(), ww.grad.float()] + ([bb.grad.float()] if mode == "plain" else [bn.weight.grad.float()]))
    ref = src.float() @ w.t() + b
    if mode == "plain":
        assert _rel_l2(res[0][0], ref) < 2e-3
    for a, c, what in zip(res[0], res[1], ("y", "dsrc", "dw", "dbias / dgamma")):
        assert _rel_l2(a, c) < 2e-3, f"{what}: dense route vs rulebook kernel {_rel_l2(a, c)}"
