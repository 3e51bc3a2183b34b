"""Dev tool: run-to-run variation of the eval forward from identical state (finds races: any difference beyond float-atomic
noise of the spectral-norm reduction is a bug)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maggie_b200.config import CfgNode
from maggie_b200.network import build_model
from oracle import make_golden as G
import synthdata as synth
case = sys.argv[1] if len(sys.argv) > 1 else "eval_128_3inst_maskos8"
kw, _ = G.CASES[case]
m, _ = build_model(CfgNode(synth.model_cfg()))
sd = synth.synth_state_dict(m.state_dict())
m.cuda().eval()
batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in synth.make_batch(**kw).items()}
z = dict(np.load(G.GOLDEN_DIR + "/" + case + ".npz"))
ref = None
for it in range(int(os.environ.get("N", "12"))):
    m.load_state_dict(sd)
    G.seed_all()
    with torch.no_grad():
        out = m(batch, mem_feat=None)
    out = {k: v.float().cpu() for k, v in out.items()}
    d8 = (out["alpha_os8"].numpy() - z["out/alpha_os8"])
    same = (out["detail_mask"].numpy() == z["out/detail_mask"]).mean()
    bad = {k: float((np.abs(out[k].numpy() - z["out/" + k]) > 1e-2).mean()) for k in ("alpha_os4", "alpha_os1", "refined_masks")}
    msg = f"run {it}: os8 max {np.abs(d8).max():.3e} mean {np.abs(d8).mean():.3e} detail agree {same:.5f} bad {bad}"
    if ref is not None:
        msg += " | vs run0: " + ", ".join(f"{k} {float((out[k] - ref[k]).abs().max()):.2e}" for k in ("alpha_os8", "alpha_os4", "alpha_os1"))
    else:
        ref = out
    print(msg, flush=True)
