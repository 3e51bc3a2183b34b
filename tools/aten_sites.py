"""Dev tool: which lines of maggie_b200 issue ATen (torch) operators during one eager C2 training step - the glue still
to be replaced by native kernels.  A TorchDispatchMode logs every dispatched op that touches a CUDA tensor together with
the innermost maggie_b200 frame on the Python stack (autograd-generated backward nodes show up as `<autograd>`)."""
import collections
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import random
import torch
from torch.utils._python_dispatch import TorchDispatchMode

from maggie_b200.config import CfgNode
from maggie_b200.network import build_model
import synthdata as synth

SKIP = {"aten::view", "aten::_unsafe_view", "aten::permute", "aten::reshape", "aten::detach", "aten::alias", "aten::expand",
        "aten::slice.Tensor", "aten::select.int", "aten::transpose.int", "aten::t", "aten::unsqueeze", "aten::squeeze.dim",
        "aten::as_strided", "aten::empty.memory_format", "aten::empty_like", "aten::empty_strided", "aten::unbind.int",
        "aten::split.Tensor", "aten::flatten.using_ints", "aten::_reshape_alias", "aten::squeeze", "aten::is_pinned",
        "aten::record_stream", "aten::lift_fresh", "aten::narrow", "aten::view.dtype", "aten::new_empty", "aten::unflatten.int"}


class Log(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.count = collections.Counter()

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = func.name() if hasattr(func, "name") else str(func)
        if name not in SKIP:
            flat = [a for a in list(args) + list((kwargs or {}).values()) if torch.is_tensor(a)]
            flat += [b for a in args if isinstance(a, (list, tuple)) for b in a if torch.is_tensor(b)]
            if any(t.is_cuda for t in flat) or not flat:
                site = "<autograd>"
                for fr in reversed(traceback.extract_stack()[:-1]):
                    if "maggie_b200" in fr.filename and "site-packages" not in fr.filename:
                        site = f"{os.path.relpath(fr.filename, ROOT)}:{fr.lineno} {fr.name}"
                        break
                self.count[(name, site)] += 1
        return func(*args, **(kwargs or {}))


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(1234)
    model, _ = build_model(CfgNode(synth.model_cfg()))
    model.to(dev).train()
    it = int(os.environ.get("ITER", "1"))
    batch = synth.make_batch(b=8, n_f=1, n_i=3, H=512, W=512, edge_px=6.0, seed=1234, train=True, it=it)
    batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items() if k not in ("fg", "bg")}

    def step():
        np.random.seed(7), random.seed(7)
        for p in model.parameters():
            p.grad = None
        _, loss = model(batch, mem_feat=None)
        (loss["total"] * 64.0).backward()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    log = Log()
    with log:
        step()
    torch.cuda.synchronize()
    total = sum(log.count.values())
    print(f"{total} ATen ops with CUDA tensors in one eager step (iter={it}); by call site:")
    for (name, site), n in sorted(log.count.items(), key=lambda kv: -kv[1]):
        print(f"{n:5d}  {name:38s} {site}")


if __name__ == "__main__":
    main()
