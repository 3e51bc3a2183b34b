"""Host-side pieces of maggie_b200.ops that are plain torch compositions (run anywhere): the GEMM-backward id embedding,
the torch path of the alpha heads, the index helpers that avoid python-list indexing."""
import torch
import torch.nn.functional as F

from maggie_b200 import ops


def test_id_embedding_matches_f_embedding_forward_and_backward():
    torch.manual_seed(0)
    table = torch.randn(11, 16, requires_grad=True)
    ids = torch.randint(0, 11, (3, 50))
    ids[0, :5] = 10
    out = ops.id_embedding(ids, table, torch.float32)
    g = torch.randn_like(out)
    out.backward(g)
    t2 = table.detach().clone().requires_grad_(True)
    ref = F.embedding(ids, t2)
    ref.backward(g)
    assert torch.equal(out, ref)
    assert torch.allclose(table.grad, t2.grad, atol=1e-5)


def test_upsample_tanh_torch_path_with_plane_scale():
    torch.manual_seed(1)
    x = torch.randn(2, 3, 8, 8)
    ps = torch.tensor([[1.0, 0.0, 1.0], [0.0, 1.0, 1.0]])
    got = ops.upsample_tanh(x, scale=4.0, plane_scale=ps)
    ref = (torch.tanh(F.interpolate(x, scale_factor=4.0, mode="bilinear", align_corners=False)) + 1.0) / 2.0 * ps[:, :, None, None]
    assert got.shape == (2, 3, 32, 32) and torch.equal(got, ref)
    assert torch.equal(ops.upsample_tanh(x), (torch.tanh(x) + 1.0) / 2.0)
    assert torch.equal(ops.upsample_tanh(x, size=(64, 64)),
                       (torch.tanh(F.interpolate(x, size=(64, 64), mode="bilinear", align_corners=False)) + 1.0) / 2.0)


def test_take_put_are_index_select_and_index_copy():
    t = torch.arange(2 * 5 * 3.0).reshape(2, 5, 3)
    assert torch.equal(ops.take(t, 1, [4, 0, 2]), t[:, [4, 0, 2]])
    out = torch.zeros(2, 7, 3)
    ops.put(out, 1, [6, 1, 3], t[:, :3])
    ref = torch.zeros(2, 7, 3)
    ref[:, [6, 1, 3]] = t[:, :3]
    assert torch.equal(out, ref)
    assert ops.index_tensor([6, 1, 3], t.device) is ops.index_tensor((6, 1, 3), t.device)   # cached


def test_token_logits_torch_path():
    torch.manual_seed(2)
    tok, x = torch.randn(2, 10, 64), torch.randn(6, 64, 4, 5)
    got = ops.token_logits(tok, x, 3)
    ref = torch.einsum("bqc,btchw->btqhw", tok, x.reshape(2, 3, 64, 4, 5)).flatten(0, 1)
    assert got.shape == (6, 10, 4, 5) and torch.allclose(got, ref)


def test_side_branch_runs_inline_without_a_gpu():
    from maggie_b200 import dense
    x = torch.ones(2, 3)
    calls = []
    y, join = dense.side_branch(x, torch.nn.BatchNorm2d(3), 5, lambda: (calls.append(1), x * 2)[1])
    assert calls == [1] and torch.equal(y, x * 2)
    assert join() is None


def test_fused_optimizer_refuses_cpu_parameters():
    import pytest
    from maggie_b200.dp import FlatGradAllReduce
    from maggie_b200.optim import FusedAdamW
    flat = FlatGradAllReduce(torch.nn.Linear(3, 2).parameters())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FusedAdamW(flat)


def test_flat_gradient_hand_over_single_process():
    """zero() drops the gradients, pack() gathers them into the flat buffer and re-points p.grad at its views."""
    from maggie_b200.dp import FlatGradAllReduce
    torch.manual_seed(3)
    lin = torch.nn.Linear(4, 3)
    frozen = torch.nn.Parameter(torch.zeros(2), requires_grad=False)
    flat = FlatGradAllReduce(list(lin.parameters()) + [frozen])
    assert flat.flat.numel() == 15 and len(flat.params) == 2
    flat.zero()
    assert all(p.grad is None for p in lin.parameters())
    lin(torch.randn(5, 4)).sum().backward()
    want = torch.cat([p.grad.flatten().clone() for p in lin.parameters()])
    assert flat.allreduce() is flat.flat              # one process: nothing is packed, nothing is reduced
    assert torch.equal(flat.pack(), want)
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(flat.params, flat.views))
    lin.bias.grad = None                               # a parameter without a gradient contributes zeros
    assert torch.equal(flat.pack()[12:], torch.zeros(3))


def test_flat_buffer_puts_the_bank_weights_first_and_keeps_the_parameter_order():
    """With the model's weight bank the conv weights own the head of the flat buffer in the bank's order (K0's grouped
    backward writes their gradients there directly); the parameter order - which numbers the optimizer state - is unchanged."""
    from types import SimpleNamespace
    from maggie_b200.dp import FlatGradAllReduce
    a, b, c, d = (torch.nn.Parameter(torch.zeros(n)) for n in (3, 5, 2, 4))
    bank = SimpleNamespace(entries=[SimpleNamespace(w=d), SimpleNamespace(w=b)], grad_target=None)
    flat = FlatGradAllReduce([a, b, c, d], bank=bank)
    assert [id(p) for p in flat.params] == [id(a), id(b), id(c), id(d)]
    assert flat.offsets == [9, 4, 12, 0]
    assert bank.grad_target.data_ptr() == flat.flat.data_ptr() and bank.grad_target.numel() == 9
    for p, g in zip((a, b, c, d), (1.0, 2.0, 3.0, 4.0)):
        p.grad = torch.full_like(p, g)
    assert flat.pack().tolist() == [4.0] * 4 + [2.0] * 5 + [1.0] * 3 + [3.0] * 2
    # a bank with a weight outside the parameter list (frozen): plain layout, no target
    bank2 = SimpleNamespace(entries=[SimpleNamespace(w=torch.nn.Parameter(torch.zeros(2), requires_grad=False))], grad_target=None)
    flat2 = FlatGradAllReduce([a, b], bank=bank2)
    assert flat2.offsets == [0, 3] and bank2.grad_target is None


def test_ellipse_width_draws_consume_the_random_stream_like_the_reference():
    """`_draw_widths` draws all slots with one vectorised call; the reference (utils/utils.py:45-50) draws them one by one:
    same values, same generator state afterwards."""
    import numpy as np
    from maggie_b200.network.decoder import _draw_widths
    for k in (15, 27, 30):
        np.random.seed(11)
        want = [int(np.random.randint(1, k)) for _ in range(80)]
        after_want = np.random.randint(0, 1 << 30)
        np.random.seed(11)
        got = _draw_widths(80, k, True)
        after_got = np.random.randint(0, 1 << 30)
        assert got == want and after_got == after_want and all(isinstance(v, int) for v in got)
    assert _draw_widths(4, 30, False) == [15] * 4
