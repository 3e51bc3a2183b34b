"""Fused optimizer tail (K14): `scaler.unscale_(opt)` + `clip_grad_norm_(params, 0.01)` + `AdamW.step()` of the reference's
training loop (engine/train.py:265-283, engine/optim.py:118) as two multi-tensor launches on the flat gradient buffer of
`maggie_b200.dp.FlatGradAllReduce`, without any host synchronisation (an inf / nan gradient skips the update on the device,
as GradScaler.step does).

    flat = FlatGradAllReduce(model.parameters())
    opt = FusedAdamW(flat, lr=1.5e-4, betas=(0.5, 0.999), weight_decay=0.01, clip_norm=0.01)
    ...
    flat.zero(); (loss * scale).backward(); flat.allreduce(); opt.step(grad_scale=scale)

`opt.param_groups[0]['lr']` is read at every step, so torch's LR schedulers drive it unchanged."""
import ctypes

import numpy as np
import torch

from . import _lib

CHUNK = 16384


class _OptimTensor(ctypes.Structure):
    _fields_ = [("param", ctypes.c_void_p), ("flat_off", ctypes.c_int64), ("numel", ctypes.c_int64)]


class FusedAdamW:
    def __init__(self, flat, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, clip_norm=0.01):
        self.flat = flat                                   # FlatGradAllReduce: parameters, flat gradient buffer, views
        params = flat.params
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdamW needs CUDA parameters; there is no CPU fallback")
        assert all(p.dtype == torch.float32 and p.is_contiguous() for p in params)
        self.param_groups = [dict(params=params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)]
        self.clip_norm = clip_norm
        n = flat.flat.numel()
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.acc = torch.zeros(2, dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.float32, device=dev)
        self.report = torch.zeros(2, dtype=torch.float32, device=dev)   # (gradient norm, found_inf) of the last step
        arr = (_OptimTensor * len(params))()
        items = []
        for i, (p, off) in enumerate(zip(params, flat.offsets)):
            arr[i].param, arr[i].flat_off, arr[i].numel = p.data_ptr(), off, p.numel()
            items += [(i, o) for o in range(0, p.numel(), CHUNK)]
        self._sig = tuple(p.data_ptr() for p in params)
        self.tensors = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        self.items = torch.from_numpy(np.asarray(items, dtype=np.int32)).to(dev)
        self.n_items = len(items)
        self._skip_key, self._skip = (), None               # device mask of the tensors without a gradient (rarely changes)

    def zero_grad(self, set_to_none=True):
        self.flat.zero()

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        """One update from the gradients currently held by the parameters (packed into the flat buffer here if the
        all-reduce has not done it already).  grad_scale: the loss scale the backward ran with."""
        if self._sig != tuple(p.data_ptr() for p in self.flat.params):
            raise RuntimeError("FusedAdamW: parameter storage moved after construction (.to() / load into new tensors)")
        grad = self.flat.pack()
        if self.flat.no_grad != self._skip_key:
            # e.g. the four `dummy_downscale` weights: trainable flag set, never part of the graph.  torch.optim.AdamW
            # skips `grad is None` parameters entirely (no weight decay, no moment decay): so does the kernel.
            self._skip_key = self.flat.no_grad
            mask = torch.zeros(len(self.flat.params), dtype=torch.uint8)
            mask[list(self._skip_key)] = 1
            self._skip = mask.to(grad.device) if self._skip_key else None
        g = self.param_groups[0]
        _lib.check(_lib.lib().mg_optim_adamw_step(
            _lib.tensor_ptr(self.tensors), _lib.tensor_ptr(self.items), self.n_items, _lib.tensor_ptr(grad), grad.numel(),
            _lib.tensor_ptr(self.m), _lib.tensor_ptr(self.v), _lib.tensor_ptr(self.acc), _lib.tensor_ptr(self.step_count),
            _lib.tensor_ptr(self.report), float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
            float(g["weight_decay"]), float(self.clip_norm), 1.0 / float(grad_scale), _lib.tensor_ptr(self._skip),
            _lib.stream_ptr()), "mg_optim_adamw_step")
        return self.report

    def state_dict(self):
        """The layout `torch.optim.AdamW.state_dict()` writes (what the reference engine saves to / resumes from
        `last_opt.pth`, engine/train.py): per-parameter `step` / `exp_avg` / `exp_avg_sq`, parameters numbered in order.
        Parameters that never received a gradient have no state entry, as in torch."""
        state = {}
        never = set(self._skip_key)
        for i, (p, off) in enumerate(zip(self.flat.params, self.flat.offsets)):
            n = p.numel()
            if i not in never:
                state[i] = dict(step=self.step_count.clone().reshape(()), exp_avg=self.m[off:off + n].view_as(p).clone(),
                                exp_avg_sq=self.v[off:off + n].view_as(p).clone())
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group.update(params=list(range(len(self.flat.params))), amsgrad=False, maximize=False, foreach=None, capturable=False,
                     differentiable=False, fused=None)
        return dict(state=state, param_groups=[group])

    def load_state_dict(self, sd):
        """Accepts a `torch.optim.AdamW` state dict over the same parameter list (or this class's own)."""
        steps = []
        for i, (p, off) in enumerate(zip(self.flat.params, self.flat.offsets)):
            n = p.numel()
            st = sd["state"].get(i)
            if st is not None:
                self.m[off:off + n].copy_(st["exp_avg"].reshape(-1)), self.v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                steps.append(float(st["step"]))
            else:
                self.m[off:off + n].zero_(), self.v[off:off + n].zero_()
        if steps:
            self.step_count.fill_(max(steps))     # one shared step counter: every parameter with state is stepped together
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in sd["param_groups"][0]:
                self.param_groups[0][k] = sd["param_groups"][0][k]
