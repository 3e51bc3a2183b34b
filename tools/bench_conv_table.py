"""Dev tool: per-layer CUDA-event timing (L2 flushed, weight packs cached) of fprop / dgrad / wgrad over the C2 layer table.
`--splitk` adds the fprop / dgrad times with the experimental split-K kernel (K2s) switched on, side by side."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from maggie_b200 import dense

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
_orig = dense.pack_weight
_memo = {}
def _cached(w, ci_pad=None):
    key = (w.data_ptr(), tuple(w.shape), tuple(w.stride()), ci_pad)
    if key not in _memo:
        _memo[key] = (_orig(w, ci_pad), w)
    return _memo[key][0]
dense.pack_weight = _cached

def timeit(fn, n=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]

tot = [0.0, 0.0, 0.0, 0.0, 0.0]
N = 8
AB = "--splitk" in sys.argv
print(f"{'layer':34s} cnt |   fwd us   TF/s |  dgrad us  TF/s |  wgrad us  TF/s | hbm-floor us (fwd)" + (" | split-K fwd / dgrad us" if AB else ""))
for (hw, ci, co, k, s, d, tr, cnt) in bench.C2_CONVS:
    x = torch.randn(N, hw, hw, ci, device=dev).half()
    w = torch.randn((ci, co, k, k) if tr else (co, ci, k, k), device=dev) / (ci * k * k) ** 0.5
    g = dense.ConvGeom("convT", 4, 2, 1, 1) if tr else dense.ConvGeom("conv", k, s, d * (k // 2) if k > 1 else 0, d)
    if not tr and k == 2:
        g = dense.ConvGeom("conv", 2, 2, 0, 1)
    y = g.fwd(x, w)
    flops = 2.0 * y.shape[0] * y.shape[1] * y.shape[2] * co * ci * (4 if tr else k * k)
    dwp = torch.zeros_like(dense.pack_weight(w.permute(1, 0, 2, 3) if tr else w, ci), dtype=torch.float32)
    class B:  # minimal bank stand-in for wgrad accumulation (no zero fill inside the timing)
        G = dwp
    tf = timeit(lambda: g.fwd(x, w))
    td = timeit(lambda: g.dgrad(y, w, x.shape))
    tw = timeit(lambda: g.wgrad(y, x, w.shape, bank=B))
    hbm = (x.numel() + y.numel()) * 2 / 6.55e12 * 1e6
    tot[0] += tf * cnt; tot[1] += td * cnt; tot[2] += tw * cnt
    name = f"{hw}^2 {ci}->{co} k{k} s{s} d{d}{' T' if tr else ''}"
    extra = ""
    if AB:
        dense.SPLITK = True
        tfs, tds = timeit(lambda: g.fwd(x, w)), timeit(lambda: g.dgrad(y, w, x.shape))
        dense.SPLITK = False
        tot[3] += tfs * cnt; tot[4] += tds * cnt
        extra = f" | {tfs:8.1f} {tds:8.1f}"
    print(f"{name:34s} {cnt:3d} | {tf:8.1f} {flops/tf/1e6:6.0f} | {td:8.1f} {flops/td/1e6:6.0f} | {tw:8.1f} {flops/tw/1e6:6.0f} | {hbm:6.1f}{extra}", flush=True)
print(f"totals per step (us): fwd {tot[0]:.0f} dgrad {tot[1]:.0f} wgrad {tot[2]:.0f}" + (f" | split-K fwd {tot[3]:.0f} dgrad {tot[4]:.0f}" if AB else ""))
