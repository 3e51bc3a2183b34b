"""Dev tool: CUDA-event timing of the K9 rulebook kernels at the C2 site counts (band-shaped active sets)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maggie_b200 import ops, sparse
import synthdata as synth
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=5):
    for _ in range(2): fn()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]
al = torch.stack([synth.soft_ellipse_alphas(1, 3, 512, 512, 6.0, seed=s)[0] for s in range(8)]).to(dev)
unk = ops.unknown_mask(al, [15] * 24)
T = ops.build_sites(unk.reshape(-1, 512, 512))
print("counts", T.counts)
for lvl, (ci, co, Tn, name) in ((0, (32, 32, 9, "OS1 SubM 3x3 32->32")), (0, (64, 32, 1, "OS1 1x1 64->32")), (2, (64, 64, 9, "OS4 SubM 3x3 64->64"))):
    N = T.counts[lvl]
    src = torch.randn(N, ci, device=dev).half()
    w = torch.randn(co, 3, 3, ci, device=dev) if Tn == 9 else torch.randn(co, ci, device=dev)
    wp = sparse.pack_fwd(w)
    tab = T.nbr[lvl] if Tn == 9 else None
    dout = torch.randn(N, co, device=dev).half()
    t_f = timeit(lambda: sparse.sparse_conv_launch(src, wp, Tn, ci, co, table=tab))
    t_w = timeit(lambda: sparse._wgrad(dout, co, src, ci, tab, Tn))
    gb = N * (Tn * ci + co) * 2 / 1e9
    print(f"{name}: N={N}  conv {t_f:7.1f} us ({gb / t_f * 1e6:6.0f} GB/s gathered+written)   wgrad {t_w:7.1f} us", flush=True)
