"""Dev tool: per-layer timing of fprop / dgrad / wgrad over the C2 layer table (weight packs cached).
Two numbers per entry: `cold` = one launch between CUDA events after an L2 flush (includes the event / launch overhead),
`burst` = average of back-to-back launches over rotating operand sets, replayed from a CUDA graph (no flush, no per-launch
event, no host launch overhead).  A trailing letter
shows the kernel the launch was routed to: h = K2b / K4b (halo), m = K2h (mid), none = generic."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from maggie_b200 import _lib, dense

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
L = _lib.lib()

def cold(fn, n=5):
    for _ in range(2):
        fn(0)
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(0); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]

def burst(fn, nsets, rounds=3):
    """back-to-back launches replayed from a CUDA graph (no host launch overhead), rotating operand sets"""
    for i in range(nsets):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(rounds):
                for i in range(nsets):
                    fn(i)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (rounds * nsets)

def route(fn):
    h0, m0, w0 = L.mg_conv_halo_launches(), L.mg_conv_mid_launches() + L.mg_conv_midt_launches(), L.mg_wgrad_halo_launches()
    fn(0)
    return "h" if (L.mg_conv_halo_launches() > h0 or L.mg_wgrad_halo_launches() > w0) else ("m" if L.mg_conv_mid_launches() + L.mg_conv_midt_launches() > m0 else " ")

tot = [0.0] * 6
N = 8
print(f"{'layer':30s} cnt | fwd cold / burst us (TF/s burst) | dgrad cold / burst | wgrad cold / burst | hbm floor us")
for (hw, ci, co, k, s, d, tr, cnt) in bench.C2_CONVS:
    w = torch.randn((ci, co, k, k) if tr else (co, ci, k, k), device=dev) / (ci * k * k) ** 0.5
    g = dense.ConvGeom("convT", 4, 2, 1, 1) if tr else dense.ConvGeom("conv", k, s, d * (k // 2) if k > 1 else 0, d)
    if not tr and k == 2:
        g = dense.ConvGeom("conv", 2, 2, 0, 1)
    x0 = torch.randn(N, hw, hw, ci, device=dev).half()
    y0 = g.fwd(x0, w)
    per_set = (x0.numel() + y0.numel()) * 2
    nsets = max(2, min(12, (300 << 20) // per_set + 1))
    xs = [x0] + [torch.randn_like(x0) for _ in range(nsets - 1)]
    ys = [y0] + [torch.randn_like(y0) for _ in range(nsets - 1)]
    flops = 2.0 * y0.shape[0] * y0.shape[1] * y0.shape[2] * co * ci * (4 if tr else k * k)
    dwp = torch.zeros_like(dense.pack_weight(w.permute(1, 0, 2, 3) if tr else w, ci), dtype=torch.float32)
    class B:
        G = dwp
    fns = (lambda i: g.fwd(xs[i], w), lambda i: g.dgrad(ys[i], w, x0.shape), lambda i: g.wgrad(ys[i], xs[i], w.shape, bank=B))
    cells = []
    for j, fn in enumerate(fns):
        r, tc, tb = route(fn), cold(fn), burst(fn, nsets)
        tot[2 * j] += tc * cnt; tot[2 * j + 1] += tb * cnt
        cells.append(f"{tc:6.1f} / {tb:6.1f}{r} ({flops / tb / 1e6:4.0f})")
    hbm = per_set / 6.55e12 * 1e6
    name = f"{hw}^2 {ci}->{co} k{k} s{s} d{d}{' T' if tr else ''}"
    print(f"{name:30s} {cnt:3d} | " + " | ".join(cells) + f" | {hbm:6.1f}", flush=True)
    del xs, ys
print(f"totals per step (us): fwd cold {tot[0]:.0f} burst {tot[1]:.0f} | dgrad cold {tot[2]:.0f} burst {tot[3]:.0f} | wgrad cold {tot[4]:.0f} burst {tot[5]:.0f}")
