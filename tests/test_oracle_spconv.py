"""Known-answer tests for the spconv restatement: fully-active identities and a
hand-computed sparse case."""
import torch
import torch.nn.functional as F

from oracle import spconv_torch as sp


def _full(B, C, H, W):
    torch.manual_seed(0)
    dense = torch.randn(B, C, H, W)
    idx = torch.stack(torch.meshgrid(torch.arange(B), torch.arange(H), torch.arange(W), indexing="ij"), -1).reshape(-1, 3).int()
    feats = dense.permute(0, 2, 3, 1).reshape(-1, C)
    return dense, sp.SparseConvTensor(feats, idx, (H, W), B)


def test_subm_equals_conv2d_when_fully_active():
    dense, x = _full(2, 5, 6, 7)
    m = sp.SubMConv2d(5, 4, 3, padding=1, bias=True)
    ref = F.conv2d(dense, m.weight.permute(0, 3, 1, 2), m.bias, padding=1)
    assert torch.allclose(m(x).dense(), ref, atol=1e-5)


def test_sparse_conv_s2_and_inverse_when_fully_active():
    dense, x = _full(1, 3, 8, 8)
    down = sp.SparseConv2d(3, 6, 3, stride=2, padding=1, bias=False, indice_key="k")
    y = down(x)
    ref = F.conv2d(dense, down.weight.permute(0, 3, 1, 2), stride=2, padding=1)
    assert torch.allclose(y.dense(), ref, atol=1e-5)
    inv = sp.SparseInverseConv2d(6, 2, 3, indice_key="k", bias=False)
    z = inv(y)
    ref_t = F.conv_transpose2d(ref, inv.weight.permute(3, 0, 1, 2), stride=2, padding=1, output_padding=1)
    assert z.spatial_shape == (8, 8) and torch.allclose(z.dense(), ref_t, atol=1e-5)


def test_subm_ignores_inactive_neighbours():
    idx = torch.tensor([[0, 1, 1], [0, 1, 2], [0, 3, 3]], dtype=torch.int32)
    feats = torch.tensor([[1.0], [10.0], [100.0]])
    m = sp.SubMConv2d(1, 1, 3, bias=False)
    with torch.no_grad():
        m.weight.copy_(torch.arange(1.0, 10.0).view(1, 3, 3, 1))
    out = m(sp.SparseConvTensor(feats, idx, (5, 5), 1)).features.flatten().tolist()
    # centre tap weight 5; (1,2) is the right neighbour (tap 6) of (1,1); (1,1) the left neighbour (tap 4) of (1,2)
    assert out == [1 * 5 + 10 * 6, 10 * 5 + 1 * 4, 100 * 5]
