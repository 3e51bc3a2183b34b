"""Dev tool: two launches each (the second one is the warm one ncu keeps) of the kernels captured with `ncu --set full` for
profiles/ (C2 shapes): K2b halo conv 512^2 32->32 (+ BN statistics), K2t transposed mid conv 64^2 128->128 (+ BN statistics,
and as a data gradient), 32^2 256->256 and 16^2 512->512, K2 generic conv 64^2 128->256 stride 2, K4 wgrad 64^2 128->128, K4b halo wgrad 512^2 32->32 and 128^2 64->64, K9b persistent sparse conv and
K9c persistent sparse wgrad on the C2 OS1 site list, K8a unknown mask, K1 mask embedding.

    ncu --set full --clock-control none --import-source on -k regex:'tcgen05|sparse_|unknown_mask|mask_embed_fwd|bn_' \\
        -o gpurun_out/r2_full python tools/ncu_kernels.py
    python tools/ncu_parse.py gpurun_out/r2_full.ncu-rep        # -> profiles/r2_ncu_full_kernels.csv, r2_ncu_traffic.json
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from maggie_b200 import dense, ops, sparse
import synthdata as synth

dev = torch.device("cuda")
torch.manual_seed(0)
def mk(N, H, W, Ci, Co, k):
    return torch.randn(N, H, W, Ci, device=dev).half(), (torch.randn(Co, Ci, k, k, device=dev) / (Ci * k * k) ** 0.5)

al = torch.stack([synth.soft_ellipse_alphas(1, 3, 512, 512, 6.0, seed=s)[0] for s in range(8)]).to(dev)
al10 = torch.stack([synth.soft_ellipse_alphas(1, 10, 512, 512, 6.0, seed=s)[0] for s in range(8)]).to(dev)
img, tab = torch.randn(8, 3, 512, 512, device=dev), torch.randn(11, 3, device=dev)
msk = (al > 0.5).float().contiguous()
for rep in range(2):
    x, w = mk(8, 512, 512, 32, 32, 3)
    g = dense.ConvGeom("conv", 3, 1, 1, 1)
    stats = dense.new_stats(32, dev)
    y = dense.conv_launch(x, dense.pack_weight(w, 32), dense.conv_taps(3, 3, 1, 1, 32), grid_hw=(512, 512), stats=stats)   # K2b
    g.wgrad(y, x, w.shape)                                                                                                 # K4b
    x, w = mk(8, 128, 128, 64, 64, 3)
    g.wgrad(torch.randn(8, 128, 128, 64, device=dev).half(), x, w.shape)                                                   # K4b, 64 channels in two halves
    for (hw, ci, co) in ((64, 128, 128), (32, 256, 256), (16, 512, 512)):
        x, w = mk(8, hw, hw, ci, co, 3)
        st = dense.new_stats(co, dev)
        y = dense.conv_launch(x, dense.pack_weight(w, ci), dense.conv_taps(3, 3, 1, 1, ci), grid_hw=(hw, hw), stats=st)    # K2t + BN sums
        if hw == 64:
            g.dgrad(y, w, x.shape)                                                                                         # K2t (dgrad)
            g.wgrad(y, x, w.shape)                                                                                         # K4
            mean, inv, gam, sums = torch.zeros(co, device=dev), torch.ones(co, device=dev), torch.ones(co, device=dev), torch.zeros(2, co, device=dev)
            from maggie_b200 import _lib
            L, P, S = _lib.lib(), _lib.tensor_ptr, _lib.stream_ptr
            yo, dx = torch.empty_like(y), torch.empty_like(y)
            L.mg_bn_apply(P(y), P(gam), P(mean), None, 0, P(yo), 8, hw, hw, co, 1, S())                                    # K3
            L.mg_bn_bwd_reduce(P(y), P(yo), P(y), P(mean), P(inv), P(sums), 8, hw, hw, co, 1, S())
            L.mg_bn_bwd_apply(P(y), P(yo), P(y), P(mean), P(inv), P(gam), P(sums), P(dx), None, 8, hw, hw, co, 1, 0, None, S())
    x, w = mk(8, 64, 64, 128, 256, 3)
    dense.ConvGeom("conv", 3, 2, 1, 1).fwd(x, w)                                                                           # K2 (stride 2)
    T = ops.build_sites(ops.unknown_mask(al, [15] * 24).reshape(-1, 512, 512))
    N = T.counts[0]
    src = torch.randn(N, 32, device=dev).half()
    w9 = torch.randn(32, 3, 3, 32, device=dev)
    sparse.sparse_conv_launch(src, sparse.pack_fwd(w9), 9, 32, 32, table=T.nbr[0])                                         # K9b
    sparse._wgrad(torch.randn(N, 32, device=dev).half(), 32, src, 32, T.nbr[0], 9)                                         # K9c
    ops.unknown_mask(al10, [15] * 80)                                                                                      # K8a
    with torch.no_grad():
        ops.mask_embed(img, msk, tab, [0, 1, 2], 32)                                                                       # K1
torch.cuda.synchronize()
print("done")
