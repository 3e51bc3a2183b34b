"""-m gpu: K6 attention cores through the C ABI against the torch restatement (fp32) on the same fp16-rounded
inputs: outputs, the attention-max statistic, and all gradients.  Tolerance: fp16 storage of the many-side
tensors -> 3e-3 of scale on outputs, 1e-2 rel-L2 on gradients."""
import pytest
import torch

import ops_ref

pytestmark = pytest.mark.gpu


def _rel_l2(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.mark.parametrize("B,F,S,pad,guid", [(2, 10, 4096, False, True), (3, 10, 1000, False, True), (2, 10, 10, True, False),
                                             (1, 16, 300, True, True), (4, 10, 4096, False, False)])
def test_few_query_attention(B, F, S, pad, guid):
    from maggie_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + S)
    E = 128
    q = torch.randn(B, F, E, generator=g).cuda()
    k = torch.randn(B, S, E, generator=g).half().cuda()
    v = torch.randn(B, S, E, generator=g).half().cuda()
    key_pad = None
    if pad:
        key_pad = (torch.rand(B, S, generator=g) > 0.7).cuda()
        key_pad[:, 0] = False
    guidance = (torch.rand(B, F, S, generator=g) > 0.5).cuda() if guid else None
    go, gs = torch.randn(B, F, E, generator=g).cuda(), torch.randn(B, F, generator=g).cuda()

    def run(fn, kk, vv):
        qq, kk, vv = q.clone().requires_grad_(True), kk.clone().requires_grad_(True), vv.clone().requires_grad_(True)
        o, st = fn(qq, kk, vv, key_pad, guidance)
        loss = (o.float() * go).sum() + ((st * gs).sum() if st is not None else 0)
        loss.backward()
        return o, st, qq.grad, kk.grad, vv.grad

    o, st, dq, dk, dv = run(ops.attention, k, v)
    orf, strf, dqr, dkr, dvr = run(ops_ref.attention, k.float(), v.float())
    assert float((o.float() - orf).abs().max()) < 3e-3 * float(orf.abs().max()) + 1e-4
    if guid:
        assert float((st - strf).abs().max()) < 1e-3
    else:
        assert st is None
    assert _rel_l2(dq, dqr) < 1e-2 and _rel_l2(dk, dkr) < 1e-2 and _rel_l2(dv, dvr) < 1e-2


@pytest.mark.parametrize("B,F,S", [(2, 10, 4096), (3, 10, 777), (1, 16, 128)])
def test_many_query_attention(B, F, S):
    from maggie_b200 import ops
    g = torch.Generator().manual_seed(S)
    E = 128
    q = torch.randn(B, S, E, generator=g).half().cuda()
    k, v = torch.randn(B, F, E, generator=g).cuda(), torch.randn(B, F, E, generator=g).cuda()
    key_pad = torch.zeros(B, F, dtype=torch.bool).cuda()
    key_pad[:, 3] = key_pad[:, F - 1] = True
    go = torch.randn(B, S, E, generator=g).half().cuda()

    def run(fn, qq):
        qq, kk, vv = qq.clone().requires_grad_(True), k.clone().requires_grad_(True), v.clone().requires_grad_(True)
        o, _ = fn(qq, kk, vv, key_pad, None)
        (o.float() * go.float()).sum().backward()
        return o, qq.grad, kk.grad, vv.grad

    o, dq, dk, dv = run(ops.attention, q)
    orf, dqr, dkr, dvr = run(ops_ref.attention, q.float())
    assert o.dtype == torch.float16
    assert float((o.float() - orf).abs().max()) < 3e-3 * float(orf.abs().max()) + 1e-4
    assert _rel_l2(dq, dqr) < 1e-2 and _rel_l2(dk, dkr) < 1e-2 and _rel_l2(dv, dvr) < 1e-2
    assert float(dk[:, 3].abs().max()) == 0 and float(dv[:, 3].abs().max()) == 0   # padded keys get no gradient
