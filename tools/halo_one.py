"""Dev tool: a few launches of the halo conv kernel (K2b) and of the generic kernel (K2) on the 512^2 32->32 layer (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maggie_b200 import dense
hw, ci, co = 512, 32, 32
x = torch.randn(8, hw, hw, ci, device="cuda").half()
w = torch.randn(co, ci, 3, 3, device="cuda") / (ci * 9) ** 0.5
wp = dense.pack_weight(w, ci)
taps = dense.conv_taps(3, 3, 1, 1, ci)
stats = dense.new_stats(co, x.device)
for _ in range(3):
    dense.conv_launch(x, wp, taps, grid_hw=(hw, hw), stats=stats)
os.environ["MAGGIE_B200_NO_HALO_CONV"] = "1"
for _ in range(2):
    dense.conv_launch(x, wp, taps, grid_hw=(hw, hw), stats=stats)
torch.cuda.synchronize()
