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
        items, off = [], 0
        for i, p in enumerate(params):
            arr[i].param, arr[i].flat_off, arr[i].numel = p.data_ptr(), off, p.numel()
            items += [(i, o) for o in range(0, p.numel(), CHUNK)]
            off += p.numel()
        self._sig = tuple(p.data_ptr() for p in params)
        self.tensors = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        self.items = torch.from_numpy(np.asarray(items, dtype=np.int32)).to(dev)
        self.n_items = len(items)

    def zero_grad(self, set_to_none=True):
        self.flat.zero()

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        """One update from the gradients currently held by the parameters (packed into the flat buffer here if the
        all-reduce has not done it already).  grad_scale: the loss scale the backward ran with."""
        if self._sig != tuple(p.data_ptr() for p in self.flat.params):
            raise RuntimeError("FusedAdamW: parameter storage moved after construction (.to() / load into new tensors)")
        grad = self.flat.pack()
        g = self.param_groups[0]
        _lib.check(_lib.lib().mg_optim_adamw_step(
            _lib.tensor_ptr(self.tensors), _lib.tensor_ptr(self.items), self.n_items, _lib.tensor_ptr(grad), grad.numel(),
            _lib.tensor_ptr(self.m), _lib.tensor_ptr(self.v), _lib.tensor_ptr(self.acc), _lib.tensor_ptr(self.step_count),
            _lib.tensor_ptr(self.report), float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
            float(g["weight_decay"]), float(self.clip_norm), 1.0 / float(grad_scale), _lib.stream_ptr()), "mg_optim_adamw_step")
        return self.report

    def state_dict(self):
        return dict(m=self.m, v=self.v, step=self.step_count, param_groups=[{k: v for k, v in self.param_groups[0].items() if k != "params"}])

    def load_state_dict(self, sd):
        self.m.copy_(sd["m"]), self.v.copy_(sd["v"]), self.step_count.copy_(sd["step"])
        self.param_groups[0].update(sd["param_groups"][0])
