"""Pure-torch CPU restatement of the slice of `spconv.pytorch` (spconv 2.x) that MaGGIe uses.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference's sparse arithmetic lives in the
third-party, un-vendored `spconv-cu120` wheel (/root/reference/requirements.txt:4, unpinned).  Call
sites restated here: maggie/network/decoder/resnet_inst_matt_spconv.py:61-130 (module construction),
:168,:184,:214,:225 (SparseConvTensor), :169,:188,:192,:232 (replace_feature), :248,:265 (dense()).

Published spconv 2.x semantics that are restated:
  * SubMConv2d      - output sites == input sites; out[p] = sum_k W[:,k,:] . in[p + k - c] over ACTIVE
                      neighbours only (implicit zero elsewhere); `padding` is irrelevant.
  * SparseConv2d    - (k=3, s=2, p=1): output site q is active iff any active input lies in its window;
                      out[q] = sum_k W[k] . in[2q - 1 + k]; the (in,out,k) pair table is stored under
                      `indice_key`.
  * SparseInverseConv2d - output sites / spatial shape are the INPUT sites of the conv that created the
                      key; out[p] = sum_{(p,q,k) in pairs} W'[k] . in[q]  (own weights, same k index).
  * weights are [Cout, kH, kW, Cin]; `.dense()` gives [batch, C, H, W] with zeros at inactive sites.
Site enumeration order is arbitrary in spconv (hash order); here outputs of SparseConv2d are sorted
lexicographically.  Everything downstream (BatchNorm1d statistics, scatter to dense) is order-free.
"""
import math

import torch
from torch import nn


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, indice_dict=None, **_):
        self.features = features
        self.indices = indices  # [N, 3] int32: (slot, y, x)
        self.spatial_shape = tuple(int(s) for s in spatial_shape)
        self.batch_size = int(batch_size)
        self.indice_dict = indice_dict if indice_dict is not None else {}

    def replace_feature(self, feature):
        return SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.indice_dict)

    def dense(self):
        H, W = self.spatial_shape
        C = self.features.shape[1]
        out = self.features.new_zeros((self.batch_size, H, W, C))
        idx = self.indices.long()
        out[idx[:, 0], idx[:, 1], idx[:, 2]] = self.features
        return out.permute(0, 3, 1, 2).contiguous()


def _index_map(indices, batch_size, H, W):
    """Dense int64 map [batch, H, W] -> row number or -1."""
    m = torch.full((batch_size, H, W), -1, dtype=torch.long, device=indices.device)
    idx = indices.long()
    m[idx[:, 0], idx[:, 1], idx[:, 2]] = torch.arange(idx.shape[0], device=indices.device)
    return m


class SparseModule(nn.Module):
    """Marker base class (only subclassed by out-of-scope baseline decoders at import time)."""


class _SparseConvBase(SparseModule):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, indice_key=None, **_):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.k = int(kernel_size)
        self.stride, self.padding = int(stride), int(padding)
        self.indice_key = indice_key
        self.weight = nn.Parameter(torch.empty(out_channels, self.k, self.k, in_channels))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            bound = 1.0 / math.sqrt(in_channels * self.k * self.k)
            self.bias = nn.Parameter(torch.empty(out_channels).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)

    def _finish(self, out):
        return out if self.bias is None else out + self.bias


class SubMConv2d(_SparseConvBase):
    def forward(self, x):
        H, W = x.spatial_shape
        feats, idx = x.features, x.indices.long()
        imap = _index_map(x.indices, x.batch_size, H, W)
        c = self.k // 2
        out = feats.new_zeros((feats.shape[0], self.out_channels))
        for ky in range(self.k):
            for kx in range(self.k):
                ny, nx = idx[:, 1] + ky - c, idx[:, 2] + kx - c
                ok = (ny >= 0) & (ny < H) & (nx >= 0) & (nx < W)
                rows = torch.full_like(ny, -1)
                rows[ok] = imap[idx[ok, 0], ny[ok], nx[ok]]
                sel = rows >= 0
                if sel.any():
                    contrib = feats[rows[sel]] @ self.weight[:, ky, kx, :].t().to(feats.dtype)
                    out = out.index_add(0, sel.nonzero(as_tuple=True)[0], contrib)
        return x.replace_feature(self._finish(out))


class SparseConv2d(_SparseConvBase):
    def forward(self, x):
        H, W = x.spatial_shape
        k, s, p = self.k, self.stride, self.padding
        Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        feats, idx = x.features, x.indices.long()
        pair_in, pair_q = [], []
        for ky in range(k):
            for kx in range(k):
                # input y = s*qy - p + ky  ->  qy = (y + p - ky) / s
                ty, tx = idx[:, 1] + p - ky, idx[:, 2] + p - kx
                ok = (ty % s == 0) & (tx % s == 0)
                qy, qx = torch.div(ty, s, rounding_mode="floor"), torch.div(tx, s, rounding_mode="floor")
                ok &= (qy >= 0) & (qy < Ho) & (qx >= 0) & (qx < Wo)
                rows = ok.nonzero(as_tuple=True)[0]
                pair_in.append(rows)
                pair_q.append((idx[rows, 0] * Ho + qy[rows]) * Wo + qx[rows])
        uniq = torch.unique(torch.cat(pair_q))  # sorted -> lexicographic (slot, y, x)
        out_indices = torch.stack([uniq // (Ho * Wo), (uniq // Wo) % Ho, uniq % Wo], dim=1).to(x.indices.dtype)
        out = feats.new_zeros((uniq.shape[0], self.out_channels))
        pairs = []
        for t, (rows, q) in enumerate(zip(pair_in, pair_q)):
            orow = torch.searchsorted(uniq, q)
            pairs.append((rows, orow))
            if rows.numel():
                ky, kx = divmod(t, k)
                out = out.index_add(0, orow, feats[rows] @ self.weight[:, ky, kx, :].t().to(feats.dtype))
        y = SparseConvTensor(self._finish(out), out_indices, (Ho, Wo), x.batch_size, x.indice_dict)
        if self.indice_key is not None:
            y.indice_dict[self.indice_key] = dict(in_indices=x.indices, in_shape=(H, W), pairs=pairs, k=k)
        return y


class SparseInverseConv2d(_SparseConvBase):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key, bias=True, **_):
        super().__init__(in_channels, out_channels, kernel_size, bias=bias, indice_key=indice_key)

    def forward(self, x):
        rec = x.indice_dict[self.indice_key]
        feats = x.features
        n_out = rec["in_indices"].shape[0]
        out = feats.new_zeros((n_out, self.out_channels))
        for t, (prow, qrow) in enumerate(rec["pairs"]):
            if prow.numel():
                ky, kx = divmod(t, rec["k"])
                out = out.index_add(0, prow, feats[qrow] @ self.weight[:, ky, kx, :].t().to(feats.dtype))
        return SparseConvTensor(self._finish(out), rec["in_indices"], rec["in_shape"], x.batch_size, x.indice_dict)


class SparseSequential(nn.Sequential):
    def forward(self, x):
        for m in self:
            if isinstance(m, _SparseConvBase):
                x = m(x)
            elif isinstance(x, SparseConvTensor):
                if x.features.shape[0] > 0:
                    x = x.replace_feature(m(x.features))
            else:
                x = m(x)
        return x
