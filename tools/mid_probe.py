"""Dev tool: K2h (mid-layer halo conv) vs the generic K2 on the mid-resolution C2 layers: error + CUDA-event timing
(L2 flushed), forward shapes and the data-gradient shapes (Ci / Co swapped)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maggie_b200 import _lib, dense
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=7):
    for _ in range(2): fn()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]
def burst(fn, n=20):
    """n back-to-back launches (no flush): per-launch time without the event / launch overhead, L2-warm operands"""
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
shapes = [(64, 128, 128), (32, 256, 256), (16, 512, 512), (64, 256, 128), (32, 512, 256)]
variants = [("generic", {"MAGGIE_B200_NO_MID_CONV": "1"}), ("mid", {}), ("mid ch32", {"MAGGIE_B200_MID_CH": "32"}),
            ("mid blk4", {"MAGGIE_B200_MID_BLOCKS": "4"}), ("mid blk3", {"MAGGIE_B200_MID_BLOCKS": "3"})]
for (hw, ci, co) in shapes:
    x = torch.randn(8, hw, hw, ci, device="cuda").half()
    w = torch.randn(co, ci, 3, 3, device="cuda") / (ci * 9) ** 0.5
    wp = dense.pack_weight(w, ci)
    taps = dense.conv_taps(3, 3, 1, 1, ci)
    flops = 2.0 * 8 * hw * hw * co * ci * 9
    stats = torch.zeros(dense.STAT_COPIES, 2, co, device="cuda")
    ref = None
    row = [f"{hw}^2 {ci}->{co}:"]
    for name, env in variants:
        for k in ("MAGGIE_B200_NO_MID_CONV", "MAGGIE_B200_MID_CH", "MAGGIE_B200_MID_BLOCKS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        m0 = _lib.lib().mg_conv_mid_launches()
        y = dense.conv_launch(x, wp, taps, grid_hw=(hw, hw))
        used = _lib.lib().mg_conv_mid_launches() > m0
        if ref is None:
            ref = y.float()
            err = ""
        else:
            err = f" err {float((y.float() - ref).abs().max()):.1e}"
        t = timeit(lambda: dense.conv_launch(x, wp, taps, grid_hw=(hw, hw)))
        tb = burst(lambda: dense.conv_launch(x, wp, taps, grid_hw=(hw, hw)))
        ts = timeit(lambda: dense.conv_launch(x, wp, taps, grid_hw=(hw, hw), stats=stats, pre_act="relu"))
        row.append(f"{name}{'*' if used else ''}: {t:6.1f} us ({flops / t / 1e6:4.0f} TF/s) burst {tb:6.1f} us ({flops / tb / 1e6:4.0f}) +stats {ts:6.1f}{err}")
    print("\n    ".join(row), flush=True)
