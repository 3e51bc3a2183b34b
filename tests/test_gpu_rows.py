"""-m gpu: K13 row kernels (LayerNorm with fused residual, column sums) against plain PyTorch fp32."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,E,with_res", [(32768, 128, True), (80, 128, True), (10094, 64, True), (7, 64, False), (0, 128, True)])
def test_layer_norm_residual_fwd_bwd(rows, E, with_res):
    from maggie_b200 import ops
    torch.manual_seed(rows + E)
    dev = torch.device("cuda")
    ln = torch.nn.LayerNorm(E).to(dev)
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.uniform_(-0.5, 0.5)
    a = torch.randn(rows, E, device=dev).half().requires_grad_(True)
    b = (torch.randn(rows, E, device=dev) * 0.5).half().requires_grad_(True) if with_res else None
    y = ops.layer_norm(a, ln, residual=b)
    assert y.dtype == torch.float16 and y.shape == a.shape
    gy = torch.randn(rows, E, device=dev).half()
    y.backward(gy)
    got = (a.grad.float(), ln.weight.grad.clone(), ln.bias.grad.clone(), b.grad.float() if with_res else None)
    a2 = a.detach().float().requires_grad_(True)
    s = a2 + b.detach().float() if with_res else a2
    if with_res:
        s = (a.detach() + b.detach()).float().detach().requires_grad_(True)   # the fp16 sum torch's `tgt + o` produces
    ln.weight.grad = ln.bias.grad = None
    ref = F.layer_norm(s, (E,), ln.weight, ln.bias, ln.eps)
    ref.backward(gy.float())
    if rows == 0:
        return
    assert (y.float() - ref).abs().max() < 4e-3
    ds = s.grad
    assert (got[0] - ds).abs().max() < 2e-2 * max(1.0, float(ds.abs().max()))
    if with_res:
        assert torch.equal(got[0], got[3])
    tol = 2e-3 * rows ** 0.5 + 1e-2
    assert (got[1] - ln.weight.grad).abs().max() < tol and (got[2] - ln.bias.grad).abs().max() < tol


@pytest.mark.parametrize("rows,C,stride", [(349968, 32, 32), (29815, 64, 64), (1000, 32, 64), (5, 128, 128), (0, 32, 32)])
def test_col_sum(rows, C, stride):
    from maggie_b200 import ops
    torch.manual_seed(rows)
    x = torch.randn(rows, stride, device="cuda").half()[:, :C]
    got = ops.col_sum(x)
    ref = x.double().sum(0)
    assert got.dtype == torch.float32 and got.shape == (C,)
    assert (got.double() - ref).abs().max() < 1e-3 * max(1.0, rows ** 0.5)


@pytest.mark.parametrize("b,n_f,Q,h,w", [(8, 1, 10, 64, 64), (2, 3, 10, 16, 24), (1, 1, 3, 9, 7)])
def test_token_logits_fwd_bwd(b, n_f, Q, h, w):
    """K13 token logits against the reference einsum('bqc,btchw->btqhw') in fp32 (and its autograd)."""
    from maggie_b200 import ops
    torch.manual_seed(b + h)
    dev = torch.device("cuda")
    tok = torch.randn(b, Q, 64, device=dev, requires_grad=True)
    x = torch.randn(b * n_f, 64, h, w, device=dev).half().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = ops.token_logits(tok, x, n_f)
    g = torch.randn_like(y)
    y.backward(g)
    tok2 = tok.detach().clone().requires_grad_(True)
    x2 = x.detach().float().clone().requires_grad_(True)
    ref = torch.einsum("bqc,btchw->btqhw", tok2, x2.reshape(b, n_f, 64, h, w)).flatten(0, 1)
    ref.backward(g)
    assert y.shape == ref.shape and (y - ref).abs().max() < 1e-4 * max(1.0, float(ref.abs().max()))
    assert (tok.grad - tok2.grad).abs().max() < 1e-3 * max(1.0, float(tok2.grad.abs().max()))
    assert (x.grad.float() - x2.grad).abs().max() < 4e-3 * max(1.0, float(x2.grad.abs().max()))   # fp16 gradient


def test_split_rows_gradient_is_one_concatenation():
    """q / k / v chunks of a packed in_proj matrix as views; the gradient comes back as the full matrix."""
    import torch.nn.functional as F
    from maggie_b200 import ops
    g = torch.Generator().manual_seed(4)
    w = torch.randn(3 * 64, 32, generator=g).cuda().requires_grad_(True)
    x = torch.randn(10, 32, generator=g).cuda()
    wq, wk, wv = ops.split_rows(w, 3)
    assert wq.data_ptr() == w.data_ptr() and wk.data_ptr() == w[64:128].data_ptr()
    (ops.small_linear(x, wq).sum() * 1.0 + ops.small_linear(x, wv, pos=x, relu=True).sum() * 2.0).backward()
    w2 = w.detach().clone().requires_grad_(True)
    (F.linear(x, w2[:64]).sum() + F.relu(F.linear(x + x, w2[128:])).sum() * 2.0).backward()
    assert torch.allclose(w.grad, w2.grad, rtol=1e-5, atol=1e-6) and float(w.grad[64:128].abs().max()) == 0.0
