"""Dev tool: alpha_os8 error vs golden for native convs, torch fp16 convs and torch fp32 convs (same host code)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import ops_ref
from maggie_b200 import ops
from maggie_b200.config import CfgNode
from maggie_b200.network import build_model
from oracle import make_golden as G
import synthdata as synth

def run(case, mode):
    kw, training = G.CASES[case]
    z = dict(np.load(os.path.join(G.GOLDEN_DIR, case + ".npz")))
    m, _ = build_model(CfgNode(synth.model_cfg()))
    m.load_state_dict(synth.synth_state_dict(m.state_dict()))
    m.cuda().train(training)
    m.decoder.inst_spec_layer.dropout.p = 0.0
    batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in synth.make_batch(**kw).items()}
    G.seed_all()
    saved = ops.conv_bn_act
    native_me = ops.mask_embed
    try:
        if mode != "native":
            ops.conv_bn_act = ops_ref.conv_bn_act
            dt = torch.float16 if mode == "torch16" else torch.float32
            ops.mask_embed = lambda *a, **k: native_me(*a, **k).to(dt)
        with torch.set_grad_enabled(training):
            out = m(batch, mem_feat=None)
            out = out[0] if training else out
    finally:
        ops.conv_bn_act, ops.mask_embed = saved, native_me
    d = {k: float(np.abs(v.detach().float().cpu().numpy() - z["out/" + k]).max()) for k, v in out.items() if k != "detail_mask"}
    d["mean8"] = float(np.abs(out["alpha_os8"].detach().float().cpu().numpy() - z["out/alpha_os8"]).mean())
    agree = float((out["detail_mask"].cpu().numpy() == z["out/detail_mask"]).mean())
    print(f"{case:28s} {mode:8s} " + " ".join(f"{k}={v:.2e}" for k, v in d.items()) + f" mask_agree={agree:.5f}")

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
for case in G.CASES:
    for mode in ("torch32", "torch16", "native"):
        run(case, mode)
