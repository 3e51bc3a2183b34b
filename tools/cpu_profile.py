"""Dev tool: cProfile of the host side of C2 training steps (graphs on), to see where the CPU time of a step goes."""
import cProfile, os, pstats, random, sys, time
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
model.enable_cuda_graphs(os.environ.get("GRAPHS", "1") == "1")
flat = FlatGradAllReduce(model.parameters(), bank=model.bank)
batch = synth.make_batch(b=8, n_f=1, n_i=3, H=512, W=512, edge_px=6.0, train=True, it=int(os.environ.get("ITER", "1")))
batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
torch.cuda.synchronize()
ev = torch.cuda.Event(); ev.record(); batch["ready_event"] = ev

def step():
    np.random.seed(7); random.seed(7)
    flat.zero()
    _, loss = model(batch, mem_feat=None)
    (loss["total"] * 128.0).backward()

for _ in range(4):
    step()
torch.cuda.synchronize()
N = 5
t0 = time.perf_counter()
for _ in range(N):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"cpu enqueue ms/step {(t1 - t0) * 1e3 / N:.2f}; with final sync {(t2 - t0) * 1e3 / N:.2f}")
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(45)
st.sort_stats("cumtime").print_stats(70)
