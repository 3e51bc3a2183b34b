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
variants = [("generic", {"MAGGIE_B200_MID_CONV": "0"}), ("mid", {}), ("mid ch32", {"MAGGIE_B200_MID_CH": "32"}),
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
        for k in ("MAGGIE_B200_MID_CONV", "MAGGIE_B200_MID_CH", "MAGGIE_B200_MID_BLOCKS"):
            os.environ.pop(k, None)
        os.environ["MAGGIE_B200_NO_MIDT_CONV"] = "1"       # this tool compares K2h (opt-in) with the generic kernel
        os.environ.setdefault("MAGGIE_B200_MID_CONV", "h")
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

# ---- timeline of one K2h launch (globaltimer stamps per CTA, see mg_conv_mid_trace)
import numpy as np
for k in ("MAGGIE_B200_MID_CONV", "MAGGIE_B200_MID_CH", "MAGGIE_B200_MID_BLOCKS"):
    os.environ.pop(k, None)
os.environ["MAGGIE_B200_MID_CONV"] = "h"
names = ["entry", "pdl_wait done", "first patch", "first weights", "MMAs issued", "accum complete", "epilogue done", "exit"]
for (hw, ci, co) in ((64, 128, 128), (32, 256, 256)):
    x = torch.randn(8, hw, hw, ci, device="cuda").half()
    w = torch.randn(co, ci, 3, 3, device="cuda") / (ci * 9) ** 0.5
    wp, taps = dense.pack_weight(w, ci), dense.conv_taps(3, 3, 1, 1, ci)
    buf = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
    for warm in (True, False):
        if not warm:
            flush.fill_(1)
        torch.cuda.synchronize()
        _lib.lib().mg_conv_mid_trace(buf.data_ptr())
        dense.conv_launch(x, wp, taps, grid_hw=(hw, hw))
        torch.cuda.synchronize()
        _lib.lib().mg_conv_mid_trace(None)
        t = buf.cpu().numpy().reshape(148, 8).astype(np.float64)
        t = t[t[:, 0] > 0]
        t0 = t[:, 0].min()
        rel = (t - t0) / 1e3
        print(f"{hw}^2 {ci}->{co} ({'L2-warm' if warm else 'L2 flushed'}): {len(t)} CTAs; us since first CTA entry: median (min .. max)")
        for j, n in enumerate(names):
            print(f"    {n:16s} {np.median(rel[:, j]):7.2f}  ({rel[:, j].min():6.2f} .. {rel[:, j].max():6.2f})")
