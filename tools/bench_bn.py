"""Dev tool: CUDA-event timing (L2 flushed) of the K3 BatchNorm kernels at C2 layer shapes, with achieved GB/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maggie_b200 import _lib
L = _lib.lib(); P = _lib.tensor_ptr; S = _lib.stream_ptr
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=5):
    for _ in range(2): fn()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]
def graph(fn, n=20):
    """n back-to-back launches replayed from a CUDA graph: L2-warm operands, no host launch cost"""
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
for (N, H, C) in ((8, 512, 32), (8, 256, 32), (8, 128, 64), (8, 64, 128), (8, 32, 256), (8, 16, 512)):
    r = torch.randn(N, H, H, C, device="cuda").half(); y = torch.relu(r); dy = torch.randn_like(r); dx = torch.empty_like(r); yo = torch.empty_like(r)
    mean = torch.zeros(C, device="cuda"); inv = torch.ones(C, device="cuda"); g = torch.ones(C, device="cuda"); sums = torch.zeros(2, C, device="cuda")
    mb = r.numel() * 2 / 1e6
    t1 = timeit(lambda: L.mg_bn_apply(P(r), P(g), P(mean), None, 0, P(yo), N, H, H, C, 1, S()))
    t2 = timeit(lambda: L.mg_bn_bwd_reduce(P(dy), P(y), P(r), P(mean), P(inv), P(sums), N, H, H, C, 1, S()))
    t3 = timeit(lambda: L.mg_bn_bwd_apply(P(dy), P(y), P(r), P(mean), P(inv), P(g), P(sums), P(dx), None, N, H, H, C, 1, 0, None, S()))
    f1 = lambda: L.mg_bn_apply(P(r), P(g), P(mean), None, 0, P(yo), N, H, H, C, 1, S())
    f2 = lambda: L.mg_bn_bwd_reduce(P(dy), P(y), P(r), P(mean), P(inv), P(sums), N, H, H, C, 1, S())
    f3 = lambda: L.mg_bn_bwd_apply(P(dy), P(y), P(r), P(mean), P(inv), P(g), P(sums), P(dx), None, N, H, H, C, 1, 0, None, S())
    print(f"    graph replay (L2-warm below ~40 MB): apply {graph(f1):6.1f} us | bwd_reduce {graph(f2):6.1f} us | bwd_apply {graph(f3):6.1f} us")
    print(f"{N}x{H}x{H}x{C} ({mb:6.1f} MB): apply {t1:6.1f} us {2*mb/t1:6.2f} TB/s | bwd_reduce {t2:6.1f} us {3*mb/t2:6.2f} TB/s | bwd_apply {t3:6.1f} us {4*mb/t3:6.2f} TB/s", flush=True)
