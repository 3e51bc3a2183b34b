"""Parameter containers with the reference's attribute paths / state-dict names.

The containers own tensors only; arithmetic goes through `maggie_b200.ops`.  BatchNorm / LayerNorm /
Embedding / MultiheadAttention / Linear are the stock torch containers (never `__call__`ed) so that
`SyncBatchNorm.convert_sync_batchnorm`, DDP wrapping and `load_state_dict(strict=True)` of reference
checkpoints keep working.
"""
import math

import torch
from torch import nn

from .. import ops


class _SNInner(nn.Module):
    """Holds weight_bar / weight_u / weight_v like `SpectralNorm(conv).module` (module/spectral_norm.py:56-71)."""

    def __init__(self, shape):
        super().__init__()
        w = torch.empty(shape)
        nn.init.xavier_uniform_(w)
        height = shape[0]
        width = w.numel() // height
        u, v = torch.randn(height), torch.randn(width)
        self.weight_u = nn.Parameter(u / (u.norm() + 1e-12), requires_grad=False)
        self.weight_v = nn.Parameter(v / (v.norm() + 1e-12), requires_grad=False)
        self.weight_bar = nn.Parameter(w)


class SNConv(nn.Module):
    """Spectral-normalised conv weight: `<name>.module.weight_{bar,u,v}`.  `weight()` performs the
    reference's one power iteration per forward (train AND eval) and returns W_bar / sigma."""

    def __init__(self, cin, cout, k, transposed=False):
        super().__init__()
        self.transposed = transposed
        self.module = _SNInner((cin, cout, k, k) if transposed else (cout, cin, k, k))

    def weight(self):
        """The normalised weight: a bank handle (K0, `maggie_b200.weights`) while a banked forward is in flight,
        otherwise a tensor from the per-layer torch composition."""
        b = getattr(self, "_bank", None)
        if b is not None and b[0].current is not None:
            return b[0].current.handle(b[1])
        m = self.module
        w = ops.spectral_weight(m.weight_bar, m.weight_u, m.weight_v)
        if getattr(self, "fold", False):  # AvgPool2d(2,2) + 1x1 conv == 2x2 stride-2 conv with W / 4 on every tap
            w = w.expand(-1, -1, 2, 2) * 0.25
        return w


class PlainConv(nn.Module):
    """Bias-free conv weight container (`nn.Conv2d` state-dict name `weight`)."""

    def __init__(self, cin, cout, k):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        nn.init.xavier_uniform_(self.weight)

    def w(self):
        """Bank handle while a banked forward is in flight, else the parameter itself."""
        b = getattr(self, "_bank", None)
        if b is not None and b[0].current is not None:
            return b[0].current.handle(b[1])
        return self.weight


class SparseConvParams(nn.Module):
    """spconv-layout weight [Cout, k, k, Cin] (+ bias), state-dict compatible with SubMConv2d /
    SparseConv2d / SparseInverseConv2d (SURVEY.md Hard part 8: layout assumed)."""

    def __init__(self, cin, cout, k, bias=False):
        super().__init__()
        self.k = k
        self.weight = nn.Parameter(torch.empty(cout, k, k, cin))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            bound = 1.0 / math.sqrt(cin * k * k)
            self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)


class Slot(nn.Module):
    """Parameter-free placeholder keeping nn.Sequential indices aligned with the reference
    (activations / pooling / upsampling modules that own no tensors)."""

    def forward(self, x):
        return x


def seq(*mods):
    return nn.Sequential(*mods)
