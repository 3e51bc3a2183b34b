"""Dev tool: timing of the halo conv kernel (K2b) on the heaviest C2 layers, with the profiling knobs of the kernel."""
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
for (hw, ci, co) in ((512, 32, 32), (256, 32, 32), (128, 64, 64), (256, 32, 64)):
    x = torch.randn(8, hw, hw, ci, device="cuda").half()
    w = torch.randn(co, ci, 3, 3, device="cuda") / (ci * 9) ** 0.5
    wp = dense.pack_weight(w, ci)
    taps = dense.conv_taps(3, 3, 1, 1, ci)
    row = [f"{hw}^2 {ci}->{co}:"]
    ref = None
    for dbg in ("off", "m0", "m0d1", "m0d5"):
        os.environ["MAGGIE_B200_HALO_DEBUG"] = "0"
        if dbg == "off":
            os.environ["MAGGIE_B200_NO_HALO_CONV"] = "1"
        else:
            os.environ["MAGGIE_B200_NO_HALO_CONV"] = "0"
            os.environ["MAGGIE_B200_HALO_MODE"] = dbg[1]
            if "d" in dbg:
                os.environ["MAGGIE_B200_HALO_DEBUG"] = dbg.split("d")[1]
        y = dense.conv_launch(x, wp, taps, grid_hw=(hw, hw))
        if ref is None:
            ref = y.float()
        elif "d" not in dbg:
            row.append(f"[err {float((y.float() - ref).abs().max()):.2e}]")
        t = timeit(lambda: dense.conv_launch(x, wp, taps, grid_hw=(hw, hw)))
        row.append(f"dbg={dbg} {t:7.1f} us")
    print("  ".join(row), flush=True)
