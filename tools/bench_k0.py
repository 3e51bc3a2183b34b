"""Dev tool: the grouped weight preparation (K0) of the full model: forward (spectral norm + fp16 packs) and backward, CUDA
events over a CUDA-graph replay (the kernels stream ~120 MB of fp32 weights: HBM-bound by design)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maggie_b200 import ops
from maggie_b200.config import CfgNode
from maggie_b200.network import build_model
import synthdata as synth
dev = torch.device("cuda:0")
torch.manual_seed(0)
model, _ = build_model(CfgNode(synth.model_cfg()))
model.to(dev).train()
bank = model.bank
def fwd():
    prep = ops.prepare_weights(bank)
    bank.release()
    return prep
for _ in range(3): fwd()
torch.cuda.synchronize()
def timeit(fn, n=10):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]
n_w = sum(e.numel for e in bank.entries)
t_f = timeit(fwd)
prep = bank._forward(True)
prep.G.normal_()
t_b = timeit(lambda: bank._backward(prep))
print(f"{len(bank.entries)} layers, {n_w / 1e6:.1f} M weights: forward (3 kernels, {n_w * 8 / 1e6:.0f} MB moved) {t_f:.1f} us = {n_w * 8 / t_f / 1e6:.2f} TB/s | "
      f"backward (2 kernels, {n_w * 16 / 1e6:.0f} MB moved) {t_b:.1f} us = {n_w * 16 / t_b / 1e6:.2f} TB/s")
