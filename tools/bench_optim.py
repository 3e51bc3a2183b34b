"""Dev tool: time the fused optimizer tail (K14) on the real parameter set against torch's unscale + clip_grad_norm_ + AdamW."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maggie_b200.config import CfgNode
from maggie_b200.network import build_model
from maggie_b200.dp import FlatGradAllReduce
from maggie_b200.optim import FusedAdamW
import synthdata as synth

dev = torch.device("cuda:0")
torch.manual_seed(1234)
model, _ = build_model(CfgNode(synth.model_cfg()))
model.to(dev).train()
flat = FlatGradAllReduce(model.parameters())
for p in flat.params:
    p.grad = torch.randn_like(p) * 1e-2
opt = FusedAdamW(flat, lr=1.5e-4, betas=(0.5, 0.999), weight_decay=0.01, clip_norm=0.01)
ref = torch.optim.AdamW(flat.params, lr=1.5e-4, betas=(0.5, 0.999), weight_decay=0.01)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def torch_tail():
    torch._foreach_mul_([p.grad for p in flat.params], 1.0 / 128.0)
    torch.nn.utils.clip_grad_norm_(flat.params, 0.01)
    ref.step()


n = flat.flat.numel()
t_f = timeit(lambda: opt.step(grad_scale=128.0))
t_t = timeit(torch_tail)
print(f"params {n / 1e6:.2f} M in {len(flat.params)} tensors")
print(f"fused tail  : {t_f * 1e3:8.1f} us/step  ({32 * n / t_f / 1e6:.0f} GB/s of the 32 B/param it must move)")
print(f"torch tail  : {t_t * 1e3:8.1f} us/step  (foreach unscale + clip_grad_norm_ + AdamW(foreach))")
