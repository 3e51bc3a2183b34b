"""Dev tool: a short run of the native kernels at C2 shapes for `ncu --set full` captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from maggie_b200 import dense

torch.manual_seed(0)
def mk(N, H, W, Ci, Co, k):
    return torch.randn(N, H, W, Ci, device="cuda").half(), (torch.randn(Co, Ci, k, k, device="cuda") / (Ci * k * k) ** 0.5)

for (N, H, W, Ci, Co, k) in [(8, 128, 128, 64, 64, 3), (8, 64, 64, 128, 128, 3), (8, 32, 32, 256, 256, 3), (8, 256, 256, 32, 32, 3)]:
    x, w = mk(N, H, W, Ci, Co, k)
    g = dense.ConvGeom("conv", k, 1, 1, 1)
    for _ in range(2):
        y = g.fwd(x, w)
        dx = g.dgrad(y, w, x.shape)
        dw = g.wgrad(y, x, w.shape)
torch.cuda.synchronize()
print("done")
