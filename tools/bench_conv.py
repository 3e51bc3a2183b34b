"""Dev tool: CUDA-event timing of the native conv kernels (fwd / dgrad / wgrad) at the C2 layer shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from maggie_b200 import dense

torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, n=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]

LAYERS = [(8, 512, 512, 16, 32, 3, 1), (8, 512, 512, 32, 32, 3, 1), (8, 256, 256, 32, 32, 3, 1), (8, 128, 128, 64, 64, 3, 1),
          (8, 64, 64, 128, 128, 3, 1), (8, 32, 32, 256, 256, 3, 1), (8, 16, 16, 512, 512, 3, 1), (8, 64, 64, 256, 128, 3, 1),
          (8, 128, 128, 64, 128, 3, 2), (8, 16, 16, 1280, 512, 1, 1)]
tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
for (N, H, W, Ci, Co, k, s) in LAYERS:
    x = torch.randn(N, H, W, Ci, device="cuda").half()
    w = torch.randn(Co, Ci, k, k, device="cuda") / (Ci * k * k) ** 0.5
    g = dense.ConvGeom("conv", k, s, k // 2, 1)
    wp = dense.pack_weight(w, Ci)
    taps = dense.conv_taps(k, k, k // 2, 1, Ci)
    Ho, Wo = g.out_hw(H, W)
    y = dense.conv_launch(x, wp, taps, stride=s, grid_hw=(Ho, Wo))
    flops = 2.0 * N * Ho * Wo * Co * Ci * k * k
    t_f = timeit(lambda: dense.conv_launch(x, wp, taps, stride=s, grid_hw=(Ho, Wo)))
    t_d = timeit(lambda: g.dgrad(y, w, x.shape))
    dwp = torch.zeros((Co, k * k * Ci), dtype=torch.float32, device="cuda")
    t_w = timeit(lambda: dense.wgrad_launch(y, x, taps, dwp, stride=s, grid_hw=(Ho, Wo)))
    print(f"{N}x{H}x{W} {Ci:4d}->{Co:4d} k{k} s{s}: fwd {t_f:7.1f} us {flops/t_f/1e6:6.1f} TF/s | dgrad(+pack) {t_d:7.1f} us {flops/t_d/1e6:6.1f} | "
          f"wgrad {t_w:7.1f} us {flops/t_w/1e6:6.1f}", flush=True)
