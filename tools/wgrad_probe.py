"""Dev tool: K4b (halo-resident wgrad) against K4 and a torch reference on the C2 small-channel layers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maggie_b200 import dense
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
for (n, hw, ci, co) in ((2, 128, 32, 32), (8, 512, 32, 32), (8, 256, 32, 32), (8, 256, 32, 64)):
    x = torch.randn(n, hw, hw, ci, device="cuda").half()
    dy = (torch.randn(n, hw, hw, co, device="cuda") * 0.1).half()
    taps = dense.conv_taps(3, 3, 1, 1, ci)
    ref = torch.nn.grad.conv2d_weight(x[:2].float().permute(0, 3, 1, 2), (co, ci, 3, 3), dy[:2].float().permute(0, 3, 1, 2), padding=1)
    row = [f"{n}x{hw}^2 {ci}->{co}:"]
    for mode in ("off", "0", "1"):
        os.environ["MAGGIE_B200_NO_HALO_CONV"] = "1" if mode == "off" else "0"
        os.environ["MAGGIE_B200_WGRAD_HALO_MODE"] = mode if mode != "off" else "0"
        dw = torch.zeros(co, 9 * ci, device="cuda")
        dense.wgrad_launch(dy[:2].contiguous(), x[:2].contiguous(), taps, dw, grid_hw=(hw, hw))
        got = dw.view(co, 3, 3, ci).permute(0, 3, 1, 2)
        err = float((got - ref).abs().max() / ref.abs().max())
        dwb = torch.zeros(co, 9 * ci, device="cuda")
        t = timeit(lambda: dense.wgrad_launch(dy, x, taps, dwb, grid_hw=(hw, hw)))
        row.append(f"mode={mode} {t:7.1f} us relerr {err:.1e}")
    print("  ".join(row), flush=True)
