"""Dev tool: one launch each of the kernels captured with `ncu --set full` for profiles/ (C2 shapes):
K2b halo conv 512^2 32->32 (+BN statistics), K2 generic conv 64^2 128->128, K4 wgrad 64^2 128->128, K9b persistent sparse
conv and K9c persistent sparse wgrad on the C2 OS1 site list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from maggie_b200 import dense, ops, sparse
from oracle import synth

dev = torch.device("cuda")
torch.manual_seed(0)
def mk(N, H, W, Ci, Co, k):
    return torch.randn(N, H, W, Ci, device=dev).half(), (torch.randn(Co, Ci, k, k, device=dev) / (Ci * k * k) ** 0.5)

for rep in range(2):
    x, w = mk(8, 512, 512, 32, 32, 3)
    wp, taps = dense.pack_weight(w, 32), dense.conv_taps(3, 3, 1, 1, 32)
    stats = dense.new_stats(32, dev)
    dense.conv_launch(x, wp, taps, grid_hw=(512, 512), stats=stats)
    x, w = mk(8, 64, 64, 128, 128, 3)
    g = dense.ConvGeom("conv", 3, 1, 1, 1)
    y = g.fwd(x, w)
    g.wgrad(y, x, w.shape)
    al = torch.stack([synth.soft_ellipse_alphas(1, 3, 512, 512, 6.0, seed=s)[0] for s in range(8)]).to(dev)
    T = ops.build_sites(ops.unknown_mask(al, [15] * 24).reshape(-1, 512, 512))
    N = T.counts[0]
    src = torch.randn(N, 32, device=dev).half()
    w9 = torch.randn(32, 3, 3, 32, device=dev)
    sparse.sparse_conv_launch(src, sparse.pack_fwd(w9), 9, 32, 32, table=T.nbr[0])
    sparse._wgrad(torch.randn(N, 32, device=dev).half(), 32, src, 32, T.nbr[0], 9)
torch.cuda.synchronize()
print("done")
