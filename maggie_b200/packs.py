"""Grouped operand packs for the rulebook / rows kernels (K9): every weight of a stage is converted to fp16 and laid out
for the forward kernel ([Cout padded to 16][T*Cin]) and for the data-gradient kernel ([Cin][T*Cout padded to 32],
taps mirrored or not) in THREE launches per forward - one multi-tensor cast into a flat buffer and two index gathers -
instead of ~8 tiny torch kernels per layer and direction (reshape / flip / permute / pad / cast / contiguous: ~250
launches per training step).  The gather indices encode the permutation + zero padding of every layer and are built once.

The packs are plain operand buffers (the weight gradients are computed natively by the kernels that consume them), so
nothing here takes part in autograd."""
import numpy as np
import torch

_ACTIVE = []


def active():
    return _ACTIVE[-1] if _ACTIVE else None


def _pad_rows(idx2d, zero):
    """[R, K] index array -> rows padded to a multiple of 16 with the index of the zero element."""
    r = idx2d.shape[0]
    if r % 16:
        idx2d = np.concatenate([idx2d, np.full((16 - r % 16, idx2d.shape[1]), zero, np.int64)], 0)
    return idx2d


class PackSet:
    """Packs of a fixed, ordered list of weight tensors (spconv layout [Cout, kh, kw, Cin] or Linear [out, in]; views such
    as the q / k / v slices of an in_proj matrix are fine).  `prepare(ws)` refreshes them from the current values."""

    def __init__(self):
        self.sig = None

    def _build(self, ws):
        dev = ws[0].device
        offs, n = [], 0
        for w in ws:
            offs.append(n)
            n += w.numel()
        zero = n                                     # one extra element that stays 0: the padding
        fwd, bwd, self.meta = [], [], {}
        fo = bo = 0
        for w, off in zip(ws, offs):
            co, ci = w.shape[0], w.shape[-1]
            T = w.numel() // (co * ci)
            base = (np.arange(w.numel(), dtype=np.int64) + off).reshape(co, T, ci)
            f = _pad_rows(base.reshape(co, T * ci), zero)
            cop = max(32, (co + 31) // 32 * 32)
            entry = {"fwd": (fo, f.shape), "cop": cop}
            fwd.append(f.reshape(-1))
            fo += f.size
            for mirror in ((False, True) if T > 1 else (False,)):
                w3 = base[:, ::-1] if mirror else base
                wt = np.transpose(w3, (2, 1, 0))     # [Cin, T, Cout]
                if cop != co:
                    wt = np.concatenate([wt, np.full((ci, T, cop - co), zero, np.int64)], 2)
                b = _pad_rows(np.ascontiguousarray(wt).reshape(ci, T * cop), zero)
                entry["bwd", mirror] = (bo, b.shape)
                bwd.append(b.reshape(-1))
                bo += b.size
            if T == 1:
                entry["bwd", True] = entry["bwd", False]
            self.meta[(w.data_ptr(), tuple(w.shape))] = entry
        self.flat = torch.zeros(n + 1, dtype=torch.float16, device=dev)
        self.dst = [self.flat[o:o + w.numel()].view(w.shape) for w, o in zip(ws, offs)]
        self.idx_fwd = torch.from_numpy(np.concatenate(fwd)).to(dev)
        self.idx_bwd = torch.from_numpy(np.concatenate(bwd)).to(dev)
        self.sig = tuple((w.data_ptr(), tuple(w.shape), tuple(w.stride())) for w in ws)

    @torch.no_grad()
    def prepare(self, ws, need_bwd=True):
        ws = [w.detach() for w in ws]
        sig = tuple((w.data_ptr(), tuple(w.shape), tuple(w.stride())) for w in ws)
        if sig != self.sig:
            self._build(ws)
        torch._foreach_copy_(self.dst, ws)           # fp32 -> fp16, all tensors in one multi-tensor launch
        self.fwd = self.flat.index_select(0, self.idx_fwd)
        self.bwd = self.flat.index_select(0, self.idx_bwd) if need_bwd else None
        return self

    def lookup(self, w):
        return self.meta.get((w.data_ptr(), tuple(w.shape)))

    def pack_fwd(self, e):
        o, shape = e["fwd"]
        return self.fwd[o:o + shape[0] * shape[1]].view(shape)

    def pack_bwd(self, e, mirror):
        o, shape = e["bwd", bool(mirror)]
        return self.bwd[o:o + shape[0] * shape[1]].view(shape), e["cop"]

    def __enter__(self):
        _ACTIVE.append(self)
        return self

    def __exit__(self, *exc):
        _ACTIVE.pop()
        return False
