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
