"""Dev tool: forward + backward of the K12 loss kernels at the C2 size (24 compact planes of 512x512) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maggie_b200 import ops
torch.manual_seed(0)
dev = torch.device("cuda")
S, H, W = 24, 512, 512
a = [torch.rand(8, 3, H, W, device=dev, requires_grad=True) for _ in range(3)]
t = torch.rand(8, 3, H, W, device=dev)
w = [(torch.rand(8, 3, H, W, device=dev) > 0.7).float() for _ in range(3)]
for _ in range(2):
    s = ops.matte_loss_sums(a[0], a[1], a[2], t, w[0], w[1], w[2])
    s.sum().backward()
torch.cuda.synchronize()
print("done")
