"""Dev tool: which stage of the eval forward varies from run to run (identical state and inputs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maggie_b200 import ops
from maggie_b200.config import CfgNode
from maggie_b200.network import build_model
from oracle import make_golden as G
import synthdata as synth
case = "eval_128_3inst_maskos8"
kw, _ = G.CASES[case]
m, _ = build_model(CfgNode(synth.model_cfg()))
sd = synth.synth_state_dict(m.state_dict())
m.cuda().eval()
batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in synth.make_batch(**kw).items()}
caps = []
def run():
    m.load_state_dict(sd)
    G.seed_all()
    cap = {}
    enc, aspp, dec = m.encoder, m.aspp, m.decoder
    h_enc = enc.register_forward_hook(lambda mod, i, o: cap.update(enc_out=o[0].float().clone(), **{f"fea{k+1}": f.float().clone() for k, f in enumerate(o[1])}))
    h_aspp = aspp.register_forward_hook(lambda mod, i, o: cap.update(aspp=o.float().clone()))
    h_l1 = dec.layer1.register_forward_hook(lambda mod, i, o: cap.update(dec_layer1=o.float().clone()))
    h_l2 = dec.layer2.register_forward_hook(lambda mod, i, o: cap.update(dec_layer2=o.float().clone()))
    h_imd = dec.refine_OS8.register_forward_hook(lambda mod, i, o: cap.update(os8_logits=o[0].float().clone(), os8_feat=o[1].float().clone(), queries=o[2].float().clone()))
    with torch.no_grad():
        out = m(batch, mem_feat=None)
    for h in (h_enc, h_aspp, h_l1, h_l2, h_imd):
        h.remove()
    cap["w_u"] = m.state_dict()["encoder.conv1.module.weight_u"].clone()
    cap["alpha_os8"] = out["alpha_os8"].float()
    return cap
a = run()
for it in range(4):
    b = run()
    print(f"run {it + 1} vs 0: " + ", ".join(f"{k} {float((a[k] - b[k]).abs().max()):.2e}/{float(a[k].abs().max()):.1e}" for k in a), flush=True)
