"""Weight bank (K0): every dense conv weight of the model prepared by ONE grouped native op per forward.

Reference behaviour covered: `SpectralNorm.forward` (module/spectral_norm.py:22-35,73-80) - one power iteration per
forward (train AND eval, u / v updated in place), W = W_bar / sigma - for the 54 spectral-normalised convs, plus the
fp16 operand packs of all 62 dense convs.  Backward: the conv weight-gradient kernels (K4) accumulate straight into
the bank's fp32 buffer G; when every conv of the step has run its backward, autograd reaches the single `_Prep` node,
which maps G through W_bar / sigma and the layout change in two launches and hands back one gradient per master weight.

Layers that run several times per forward with their own power iteration each time (the video model's
temporal-difference head) stay outside the bank (`bankable = False`) and use `ops.spectral_weight`.
"""
import ctypes
from ctypes import c_int32, c_int64, c_void_p

import torch

from . import _lib

_ptr, _stream = _lib.tensor_ptr, _lib.stream_ptr


class WprepLayer(ctypes.Structure):
    """Mirror of `mg_wprep_layer` (include/maggie_b200.h)."""
    _fields_ = [("w", c_void_p), ("u", c_void_p), ("v", c_void_p),
                ("Co", c_int32), ("Ci", c_int32), ("taps", c_int32), ("transposed", c_int32), ("fold", c_int32),
                ("ci_pad", c_int32), ("vec_off", c_int32), ("pad_", c_int32),
                ("p_off", c_int64), ("d_off", c_int64), ("g_off", c_int64), ("grad_off", c_int64)]


def _pad16(c):
    return (c + 15) // 16 * 16


def _align(n, a=128):
    return (n + a - 1) // a * a


class _Entry:
    __slots__ = ("module", "w", "u", "v", "transposed", "fold", "Co", "Ci", "kh", "kw", "taps", "taps_out", "ci_pad",
                 "p_off", "d_off", "vec_off", "grad_off", "numel")


class BankedWeight:
    """Handle a conv op receives instead of a weight tensor: views of the prepared packs of one layer."""
    __slots__ = ("prep", "e")

    def __init__(self, prep, entry):
        self.prep, self.e = prep, entry

    @property
    def token(self):
        return self.prep.token

    @property
    def logical(self):
        """(Co, Ci, kh, kw) of the convolution actually executed (2x2 for the folded avg-pool skip)."""
        e = self.e
        return (e.Co, e.Ci, 2, 2) if e.fold else (e.Co, e.Ci, e.kh, e.kw)

    @property
    def transposed(self):
        return bool(self.e.transposed)

    @property
    def ci_pad(self):
        return self.e.ci_pad

    def _view(self, buf, off, rows, cols):
        return buf[off:off + rows * cols].view(rows, cols)

    @property
    def P(self):   # fp16 [Co, taps*ci_pad]
        e = self.e
        return self._view(self.prep.P, e.p_off, e.Co, e.taps_out * e.ci_pad)

    @property
    def D(self):   # fp16 [ci_pad, taps*Co]
        e = self.e
        return self._view(self.prep.D, e.d_off, e.ci_pad, e.taps_out * e.Co)

    @property
    def G(self):   # fp32 [Co, taps*ci_pad], zeroed at prepare time, accumulated by the wgrad kernel
        e = self.e
        return self._view(self.prep.G, e.p_off, e.Co, e.taps_out * e.ci_pad)


class Prepared:
    __slots__ = ("bank", "P", "D", "G", "vec", "scal", "token", "pools", "aux")

    def handle(self, idx):
        return BankedWeight(self, self.bank.entries[idx])


class _Prep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bank, need_bwd, *ws):
        prep = bank._forward(need_bwd)
        ctx.bank, ctx.prep = bank, prep
        prep.token = torch.zeros((), dtype=torch.float32, device=ws[0].device)
        bank._out = prep
        return prep.token

    @staticmethod
    def backward(ctx, _g):
        grads = ctx.bank._backward(ctx.prep)
        need = ctx.needs_input_grad[2:]
        return (None, None, *[g if n else None for g, n in zip(grads, need)])


class WeightBank:
    def __init__(self):
        self.entries = []
        self.current = None      # the Prepared of the forward in flight (set by the dense stage)
        self._built_for = None
        self._out = None
        self.grad_target = None  # fp32 [grad_size] owned by dp.FlatGradAllReduce: where `_backward` puts the weight gradients

    # ---------------------------------------------------------------------------------------- registration
    def attach(self, root):
        """Registers every bankable SNConv / PlainConv below `root` (in module order) and leaves a back reference
        (`_bank`) on each so that `weight()` can hand out handles while a forward is in flight."""
        from .network.layers import PlainConv, SNConv

        for m in root.modules():
            if not isinstance(m, (SNConv, PlainConv)) or not getattr(m, "bankable", True) or getattr(m, "_bank", None):
                continue
            e = _Entry()
            e.module = m
            if isinstance(m, SNConv):
                e.w, e.u, e.v = m.module.weight_bar, m.module.weight_u, m.module.weight_v
                e.transposed = int(m.transposed)
            else:
                e.w, e.u, e.v, e.transposed = m.weight, None, None, 0
            e.fold = int(getattr(m, "fold", False))
            d0, d1, e.kh, e.kw = e.w.shape
            e.Co, e.Ci = (d1, d0) if e.transposed else (d0, d1)
            e.taps = e.kh * e.kw
            e.taps_out = 4 if e.fold else e.taps
            assert e.taps <= 16 and (not e.fold or e.taps == 1) and e.Co % 16 == 0, (e.w.shape, e.fold)
            e.ci_pad = max(_pad16(e.Ci), int(getattr(m, "ci_pad_min", 0)))   # layers fed by a wider-padded tensor
            e.numel = e.w.numel()
            m._bank = (self, len(self.entries))
            self.entries.append(e)
        self._built_for = None
        return self

    # ---------------------------------------------------------------------------------------- device tables
    def _signature(self):
        return tuple((e.w.data_ptr(), e.u.data_ptr() if e.u is not None else 0) for e in self.entries)

    def _build(self):
        dev = self.entries[0].w.device
        if dev.type != "cuda":
            raise RuntimeError("maggie_b200 weight bank needs CUDA parameters; there is no CPU fallback")
        arr = (WprepLayer * len(self.entries))()
        p = d = vec = grad = 0
        vt, uu, tiles = [], [], []
        for i, e in enumerate(self.entries):
            assert e.w.dtype == torch.float32 and e.w.is_contiguous()
            e.p_off, e.d_off, e.vec_off, e.grad_off = p, d, vec, grad
            L = arr[i]
            L.w = e.w.data_ptr()
            L.u = e.u.data_ptr() if e.u is not None else None
            L.v = e.v.data_ptr() if e.v is not None else None
            L.Co, L.Ci, L.taps, L.transposed, L.fold, L.ci_pad = e.Co, e.Ci, e.taps, e.transposed, e.fold, e.ci_pad
            L.vec_off, L.p_off, L.d_off, L.g_off, L.grad_off = vec, p, d, p, grad
            d0 = e.Ci if e.transposed else e.Co
            width = (e.Co if e.transposed else e.Ci) * e.taps
            if e.u is not None:
                vt += [(i, 0, c, 0) for c in range(0, width, 64)]   # one CTA per 64 columns, all rows (deterministic)
                uu += [(i, r, 0, 0) for r in range(0, d0, 8)]
                vec = _align(vec + width + d0, 4)   # 16-byte aligned per-layer vectors (float4 paths of K0)
            ra, rb = (e.ci_pad, e.Co) if e.transposed else (e.Co, e.ci_pad)
            first = 1
            for a in range(0, ra, 16):
                for b in range(0, rb, 32):
                    tiles.append((i, a, b, first))
                    first = 0
            p += _align(e.Co * e.taps_out * e.ci_pad)
            d += _align(e.ci_pad * e.taps_out * e.Co)
            grad += e.numel
        self.layers_dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        mk = lambda rows: torch.tensor(rows if rows else [(0, 0, 0, 0)], dtype=torch.int32).to(dev)
        self.items_vt, self.items_u, self.items_tile = mk(vt), mk(uu), mk(tiles)
        self.n_vt, self.n_u, self.n_tile = len(vt), len(uu), len(tiles)
        self.p_size, self.d_size, self.vec_size, self.grad_size = p, d, max(vec, 1), grad
        self.device = dev
        self._built_for = self._signature()

    # ---------------------------------------------------------------------------------------- forward / backward
    def prepare(self):
        """Runs the grouped weight preparation and returns the `Prepared` set (also kept in `self.current` until
        `release()`).  Differentiable w.r.t. every master weight through the returned token."""
        if not self.entries:
            return None
        if self._built_for is None or self._built_for != self._signature():
            self._build()
        need_bwd = torch.is_grad_enabled() and any(e.w.requires_grad for e in self.entries)
        _Prep.apply(self, need_bwd, *[e.w for e in self.entries])
        self.current, self._out = self._out, None
        return self.current

    def release(self):
        self.current = None

    def _forward(self, need_bwd):
        dev = self.device
        prep = Prepared()
        prep.bank = self
        prep.P = torch.empty(self.p_size, dtype=torch.float16, device=dev)
        prep.D = torch.empty(self.d_size, dtype=torch.float16, device=dev) if need_bwd else None
        prep.G = torch.zeros(self.p_size, dtype=torch.float32, device=dev) if need_bwd else None
        prep.vec = torch.zeros(self.vec_size, dtype=torch.float32, device=dev)
        prep.scal = torch.empty((len(self.entries), 4), dtype=torch.float32, device=dev)
        prep.pools = None
        prep.aux = None
        _lib.check(_lib.lib().mg_wprep_fwd(_ptr(self.layers_dev), _ptr(self.items_vt), self.n_vt, _ptr(self.items_u), self.n_u,
                                           _ptr(self.items_tile), self.n_tile, _ptr(prep.vec), _ptr(prep.scal), _ptr(prep.P),
                                           _ptr(prep.D), _stream()), "mg_wprep_fwd")
        return prep

    def _backward(self, prep):
        if prep.aux is not None:   # weight gradients were accumulated on the auxiliary stream (dense.wgrad_async)
            torch.cuda.current_stream(self.device).wait_stream(prep.aux)
            prep.aux = None
        # straight into the head of the data-parallel flat buffer when there is one and no earlier gradient is alive in
        # it (gradient accumulation over several backwards keeps torch's semantics through a fresh buffer)
        tgt = self.grad_target
        if (tgt is not None and tgt.numel() == self.grad_size and tgt.device == self.device
                and all(e.w.grad is None for e in self.entries)):
            grad = tgt
        else:
            grad = torch.empty(self.grad_size, dtype=torch.float32, device=self.device)
        prep.scal[:, 3].zero_()  # <G, W_bar> accumulators (a second backward through the same graph starts clean)
        _lib.check(_lib.lib().mg_wprep_bwd(_ptr(self.layers_dev), _ptr(self.items_tile), self.n_tile, _ptr(prep.G), _ptr(prep.vec),
                                           _ptr(prep.scal), _ptr(grad), _stream()), "mg_wprep_bwd")
        return [grad[e.grad_off:e.grad_off + e.numel].view(e.w.shape) for e in self.entries]
