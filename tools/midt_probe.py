"""Dev tool: K2t (transposed mid conv) vs K2h vs the generic K2 on the mid-resolution C2 layers: error + CUDA-event timing
(L2 flushed) and back-to-back burst timing; optional slab heights (MAGGIE_B200_MIDT_ROWS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maggie_b200 import _lib, dense
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ENV = ("MAGGIE_B200_MID_CONV", "MAGGIE_B200_NO_MIDT_CONV", "MAGGIE_B200_MIDT_ROWS")
def timeit(fn, n=9):
    for _ in range(2): fn()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]
def burst(fn, n=20):
    """n back-to-back launches replayed from a CUDA graph (no host launch cost, L2-warm operands, PDL overlap)"""
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
shapes = [(64, 128, 128), (32, 256, 256), (16, 512, 512), (64, 256, 128), (32, 512, 256), (64, 128, 256), (32, 256, 512)]
rows = {64: (2, 4, 7), 32: (2, 4, 8), 16: (4, 8, 16)}
for (hw, ci, co) in shapes:
    variants = [("generic", {"MAGGIE_B200_NO_MIDT_CONV": "1"}), ("K2h", {"MAGGIE_B200_NO_MIDT_CONV": "1", "MAGGIE_B200_MID_CONV": "h"}),
                ("K2t", {})] + [(f"K2t R={r}", {"MAGGIE_B200_MIDT_ROWS": str(r)}) for r in rows[hw]]
    x = torch.randn(8, hw, hw, ci, device="cuda").half()
    w = torch.randn(co, ci, 3, 3, device="cuda") / (ci * 9) ** 0.5
    wp = dense.pack_weight(w, ci)
    taps = dense.conv_taps(3, 3, 1, 1, ci)
    flops = 2.0 * 8 * hw * hw * co * ci * 9
    stats = torch.zeros(dense.STAT_COPIES, 2, co, device="cuda")
    ref = None
    row = [f"{hw}^2 {ci}->{co}:"]
    for name, env in variants:
        for k in ENV:
            os.environ.pop(k, None)
        os.environ.update(env)
        t0 = _lib.lib().mg_conv_midt_launches()
        y = dense.conv_launch(x, wp, taps, grid_hw=(hw, hw))
        used = _lib.lib().mg_conv_midt_launches() > t0
        if ref is None:
            ref = y.float()
            err = ""
        else:
            err = f" err {float((y.float() - ref).abs().max()):.1e}"
        yo = torch.empty_like(y)
        t = timeit(lambda: dense.conv_launch(x, wp, taps, grid_hw=(hw, hw), out=yo))
        tb = burst(lambda: dense.conv_launch(x, wp, taps, grid_hw=(hw, hw), out=yo))
        ts = burst(lambda: dense.conv_launch(x, wp, taps, grid_hw=(hw, hw), stats=stats, pre_act="relu", out=yo))
        row.append(f"{name}{'*' if used else ''}: flushed {t:6.1f} us ({flops / t / 1e6:4.0f} TF/s) graph {tb:6.1f} us ({flops / tb / 1e6:4.0f}) graph+stats {ts:6.1f} ({flops / ts / 1e6:4.0f}){err}")
    print("\n    ".join(row), flush=True)

# ---- timeline of one K2t launch (globaltimer stamps per CTA, see mg_conv_midt_trace)
import numpy as np
names = ["entry", "setup+pdl_wait", "first patch", "first weights", "MMAs issued", "accum complete", "epilogue staged", "exit"]
for (hw, ci, co, r) in ((64, 128, 128, 4), (32, 256, 256, 4), (16, 512, 512, 4)):
    for k in ENV:
        os.environ.pop(k, None)
    os.environ["MAGGIE_B200_MIDT_ROWS"] = str(r)
    x = torch.randn(8, hw, hw, ci, device="cuda").half()
    w = torch.randn(co, ci, 3, 3, device="cuda") / (ci * 9) ** 0.5
    wp, taps = dense.pack_weight(w, ci), dense.conv_taps(3, 3, 1, 1, ci)
    buf = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
    dense.conv_launch(x, wp, taps, grid_hw=(hw, hw))
    torch.cuda.synchronize()
    _lib.lib().mg_conv_midt_trace(buf.data_ptr())
    dense.conv_launch(x, wp, taps, grid_hw=(hw, hw))
    torch.cuda.synchronize()
    _lib.lib().mg_conv_midt_trace(None)
    t = buf.cpu().numpy().reshape(148, 8).astype(np.float64)
    t = t[t[:, 0] > 0]
    rel = (t - t[:, 0].min()) / 1e3
    print(f"{hw}^2 {ci}->{co} R={r} (L2-warm): {len(t)} CTAs; us since first CTA entry: median (min .. max)")
    for j, n in enumerate(names):
        print(f"    {n:16s} {np.median(rel[:, j]):7.2f}  ({rel[:, j].min():6.2f} .. {rel[:, j].max():6.2f})")
