"""Dev tool: error distribution of the refined alphas (OS4 / OS1 / fused) against the reference goldens per precision mode:
max, mean, fraction of pixels beyond 1e-3 / 1e-2, and the same restricted to the refinement region."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from maggie_b200.config import CfgNode
from maggie_b200.network import build_model
from oracle import make_golden as G
import synthdata as synth

for case in [c for c in G.CASES if c.startswith("eval")]:
    kw, _ = G.CASES[case]
    z = dict(np.load(os.path.join(G.GOLDEN_DIR, case + ".npz")))
    for mode in ("fp16", "high"):
        m, _ = build_model(CfgNode(synth.model_cfg()))
        m.load_state_dict(synth.synth_state_dict(m.state_dict()))
        m.cuda().eval()
        m.set_precision(mode)
        batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in synth.make_batch(**kw).items()}
        G.seed_all()
        with torch.no_grad():
            out = m(batch, mem_feat=None)
        dm = z["out/detail_mask"] > 0
        for k in ("alpha_os4", "alpha_os1", "refined_masks"):
            e = np.abs(out[k].float().cpu().numpy() - z["out/" + k])
            ein = e[dm] if dm.any() else e
            print(f"{case:24s} {mode:5s} {k:14s} max {e.max():.2e} mean {e.mean():.2e} | in detail region ({dm.mean():.3f} of px): "
                  f"mean {ein.mean():.2e} median {np.median(ein):.2e} frac>1e-3 {np.mean(ein > 1e-3):.3f} frac>1e-2 {np.mean(ein > 1e-2):.3f} frac>1e-1 {np.mean(ein > 1e-1):.4f}")

# ---- where are the large OS1 errors?  (fp32-accurate mode, one case)
case = "eval_192x256_3inst"
kw, _ = G.CASES[case]
z = dict(np.load(os.path.join(G.GOLDEN_DIR, case + ".npz")))
m, _ = build_model(CfgNode(synth.model_cfg()))
m.load_state_dict(synth.synth_state_dict(m.state_dict()))
m.cuda().eval()
m.set_precision("high")
batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in synth.make_batch(**kw).items()}
G.seed_all()
with torch.no_grad():
    out = m(batch, mem_feat=None)
a1, r1 = out["alpha_os1"].float().cpu().numpy(), z["out/alpha_os1"]
e = np.abs(a1 - r1)
idx = np.argsort(e.ravel())[::-1][:12]
print("largest alpha_os1 errors (index, ours, reference, detail mask ours / ref, alpha_os8 ref):")
for i in idx:
    pos = np.unravel_index(i, e.shape)
    print("  ", pos, f"{a1[pos]:.4f} {r1[pos]:.4f}", int(out['detail_mask'].cpu().numpy()[pos]), int(z['out/detail_mask'][pos]), f"{z['out/alpha_os8'][pos]:.4f}")

# ---- refined-alpha error away from detail-mask disagreements (threshold flips of alpha_os8 at 1/255, 254/255)
import torch.nn.functional as F
for case in [c for c in G.CASES if c.startswith("eval")]:
    kw, _ = G.CASES[case]
    z = dict(np.load(os.path.join(G.GOLDEN_DIR, case + ".npz")))
    m, _ = build_model(CfgNode(synth.model_cfg()))
    m.load_state_dict(synth.synth_state_dict(m.state_dict()))
    m.cuda().eval()
    m.set_precision("high")
    batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in synth.make_batch(**kw).items()}
    G.seed_all()
    with torch.no_grad():
        out = m(batch, mem_feat=None)
    dis = torch.from_numpy((out["detail_mask"].cpu().numpy() != z["out/detail_mask"]).astype(np.float32))
    sh = dis.shape
    for d in (0, 2, 4, 8, 12, 16):
        near = F.max_pool2d(dis.reshape(-1, 1, *sh[-2:]), 2 * d + 1, 1, d).reshape(sh).numpy() > 0 if d else dis.numpy() > 0
        row = [f"d={d:2d} excluded {near.mean():.5f}"]
        for k in ("alpha_os4", "alpha_os1", "refined_masks"):
            e = np.abs(out[k].float().cpu().numpy() - z["out/" + k])
            row.append(f"{k} max {e[~near].max():.2e}")
        print(case, int(dis.sum()), "flipped px |", " | ".join(row))
