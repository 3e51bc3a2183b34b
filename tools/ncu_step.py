"""Dev tool: ONE C2 training step inside a cudaProfilerStart/Stop range, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/ncu_step.py
(the launch list of a single step; warm-up steps run outside the range). GRAPHS=1 replays the dense stage as CUDA graphs."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maggie_b200.config import CfgNode
from maggie_b200.network import build_model
from maggie_b200.dp import FlatGradAllReduce
import synthdata as synth

dev = torch.device("cuda:0")
torch.manual_seed(1234)
model, _ = build_model(CfgNode(synth.model_cfg()))
model.to(dev).train()
model.enable_cuda_graphs(os.environ.get("GRAPHS", "0") == "1")
flat = FlatGradAllReduce(model.parameters())
batch = synth.make_batch(b=int(os.environ.get("B", "8")), n_f=1, n_i=3, H=512, W=512, edge_px=6.0, train=True, it=1)
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
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
