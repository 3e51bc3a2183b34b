"""Dev tool: torch.profiler breakdown of one C2 training step (top CUDA kernels, CPU vs GPU time)."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from maggie_b200.config import CfgNode
from maggie_b200.network import build_model
from maggie_b200.dp import FlatGradAllReduce
import synthdata as synth

dev = torch.device("cuda:0")
torch.manual_seed(1234)
model, _ = build_model(CfgNode(synth.model_cfg()))
model.to(dev).train()
model.enable_cuda_graphs(os.environ.get("GRAPHS", "0") == "1")
flat = FlatGradAllReduce(model.parameters(), bank=model.bank)
b = int(os.environ.get("B", "8"))
batch = synth.make_batch(b=b, n_f=1, n_i=3, H=512, W=512, edge_px=6.0, train=True, it=1)
batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}

def step():
    np.random.seed(7); random.seed(7)
    flat.zero()
    _, loss = model(batch, mem_feat=None)
    (loss["total"] * 128.0).backward()
    return loss

for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter(); step(); torch.cuda.synchronize(); print("wall ms/step", (time.perf_counter() - t0) * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ka = prof.key_averages()
tot = sum(e.device_time_total for e in ka if e.device_type == torch.autograd.DeviceType.CUDA) if False else sum(e.self_device_time_total for e in ka)
print("total self CUDA ms", tot / 1e3, "kernel-ish events", sum(e.count for e in ka if e.self_device_time_total > 0))
print(ka.table(sort_by="self_cuda_time_total", row_limit=int(os.environ.get("ROWS", "45")), max_name_column_width=70))
if os.environ.get("CPU_TABLE", "0") == "1":
    print(ka.table(sort_by="self_cpu_time_total", row_limit=60, max_name_column_width=70))
if os.environ.get("STACKS", "0") == "1":
    # where the torch glue comes from: ATen ops with device time, grouped by the innermost maggie_b200 frame
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof2:
        step(); torch.cuda.synchronize()
    sites = {}
    for e in prof2.key_averages(group_by_stack_n=12):
        if not e.key.startswith("aten::") or e.self_device_time_total <= 0:
            continue
        frame = next((f for f in e.stack if "maggie_b200/" in f), "(autograd engine / other)")
        frame = frame.split("maggie_b200/")[-1]
        k = (frame, e.key)
        c, t = sites.get(k, (0, 0.0))
        sites[k] = (c + e.count, t + e.self_device_time_total)
    print("\nATen device time by call site (us, launches):")
    for (frame, op), (c, t) in sorted(sites.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get("ROWS", "45"))]:
        print(f"{t:9.1f} {c:5d}  {op:32s} {frame}")
