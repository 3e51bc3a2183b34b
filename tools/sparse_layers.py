"""Dev tool: every rulebook-kernel launch (K9 / K9b / K9c) of one C2 training step with its shape, then CUDA-graph-replay
(L2-warm) and flushed timings per distinct shape."""
import os, sys, random, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maggie_b200 import sparse, _lib
from maggie_b200.config import CfgNode
from maggie_b200.network import build_model
import synthdata as synth

dev = torch.device("cuda:0")
torch.manual_seed(1234)
model, _ = build_model(CfgNode(synth.model_cfg()))
model.to(dev).train()
batch = synth.make_batch(b=8, n_f=1, n_i=3, H=512, W=512, edge_px=6.0, train=True, it=1)
batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
def step():
    np.random.seed(7); random.seed(7)
    for p in model.parameters(): p.grad = None
    _, loss = model(batch, mem_feat=None)
    (loss["total"] * 128.0).backward()
for _ in range(2): step()
log, keep = [], {}
o_conv, o_wg = sparse.sparse_conv_launch, sparse._wgrad
def conv(src, wp, T, cin, cout, **kw):
    No = kw.get("n_out") or (kw["table"].shape[0] if kw.get("table") is not None else src.shape[0])
    key = ("conv", No, src.shape[0], cin, cout, T, kw.get("stats") is not None, kw.get("head") is not None)
    log.append(key)
    keep.setdefault(key, (src, wp, T, cin, cout, dict(kw)))
    return o_conv(src, wp, T, cin, cout, **kw)
def wg(dout, cout, src, cin, table, T):
    key = ("wgrad", dout.shape[0], src.shape[0], cin, cout, T, False, False)
    log.append(key)
    keep.setdefault(key, (dout, cout, src, cin, table, T))
    return o_wg(dout, cout, src, cin, table, T)
sparse.sparse_conv_launch, sparse._wgrad = conv, wg
step(); torch.cuda.synchronize()
sparse.sparse_conv_launch, sparse._wgrad = o_conv, o_wg
cnt = collections.Counter(log)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def graph(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / n)
    return sorted(ts)[2]
def cold(fn, n=5):
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]
tot_w = tot_c = 0.0
print(f"{'kind':6s} {'No':>8s} {'Ns':>8s} cin cout  T stats head  cnt | warm us  cold us")
for (kind, No, Ns, cin, cout, T, st, head), c in sorted(cnt.items(), key=lambda kv: (-kv[0][1], kv[0])):
    args = keep[(kind, No, Ns, cin, cout, T, st, head)]
    if kind == "conv":
        src, wp, T_, cin_, cout_, kw = args
        kw = dict(kw)
        if kw.get("out") is None and not head:
            kw["out"] = torch.empty((No, cout), dtype=torch.float16, device=dev)
        fn = lambda: o_conv(src, wp, T_, cin_, cout_, **kw)
    else:
        fn = lambda: o_wg(*args)
    try:
        tw, tc = graph(fn), cold(fn)
    except Exception as e:
        print(kind, No, Ns, cin, cout, T, "failed:", str(e)[:80]); continue
    tot_w += tw * c; tot_c += tc * c
    print(f"{kind:6s} {No:8d} {Ns:8d} {cin:3d} {cout:4d} {T:2d} {int(st):5d} {int(head):4d} {c:4d} | {tw:7.1f}  {tc:7.1f}", flush=True)
print(f"per step: warm {tot_w:.0f} us, cold {tot_c:.0f} us")
