"""Dev tool (CPU): which roundings of the 16-bit path cost how much alpha_os8 accuracy.

Runs the host model with the torch reference ops of tests/ops_ref.py (fp32) and emulates fp16 storage / fp16 MMA operands
at selected places (`.half().float()`), then compares alpha_os8 with the reference golden.  Used to design the
`precision="high"` evaluation mode (split-fp16 operands on the dense convolutions).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.nn.functional as F

import ops_ref
from maggie_b200 import ops
from maggie_b200.config import CfgNode
from maggie_b200.network import build_model
from oracle import make_golden as G
import synthdata as synth

q16 = lambda t: t.half().float()


def split2(t):
    hi = t.half().float()
    lo = (t - hi).half().float()
    return hi + lo


QUANT = {"none": lambda t: t, "fp16": q16, "split": split2}


def make_conv(qx, qw, qo):
    def conv_bn_act(x, w, bn, training, *, residual=None, **kw):
        x = QUANT[qx](x)
        w = QUANT[qw](w)
        if residual is not None:
            residual = QUANT[qo](residual)
        return QUANT[qo](ops_ref.conv_bn_act(x, w, bn, training, residual=residual, **kw))
    return conv_bn_act


def run(case, conv=("none", "none", "none"), rows="none", input_q="fp16", sparse="none"):
    kw, training = G.CASES[case]
    z = dict(np.load(os.path.join(G.GOLDEN_DIR, case + ".npz")))
    m, _ = build_model(CfgNode(synth.model_cfg()))
    m.load_state_dict(synth.synth_state_dict(m.state_dict()))
    m.train(training)
    G.seed_all()
    with ops_ref.injected():
        base_me = ops.mask_embed
        ops.conv_bn_act = make_conv(*conv)
        ops.mask_embed = lambda *a, **k: QUANT[input_q](base_me(*a, **k).float())
        saved = (ops.linear, ops.linear_rows, ops.layer_norm, ops.rows_conv)
        lin0, ln0, rc0 = ops.linear, ops.layer_norm, ops.rows_conv
        R = QUANT[rows]
        ops.linear = lambda x, w, b=None: R(lin0(R(x), R(w), b))
        ops.linear_rows = ops.linear
        ops.layer_norm = lambda x, ln, residual=None: R(ln0(R(x), ln, None if residual is None else R(residual)))
        S = QUANT[sparse]
        ops.rows_conv = lambda src, w, bias=None, **k: S(rc0(S(src), S(w), bias, **k))
        try:
            with torch.no_grad():
                out = m(synth.make_batch(**kw), mem_feat=None)
        finally:
            ops.linear, ops.linear_rows, ops.layer_norm, ops.rows_conv = saved
    d8 = np.abs(out["alpha_os8"].numpy() - z["out/alpha_os8"])
    same = out["detail_mask"].numpy() == z["out/detail_mask"]
    dr = np.abs(out["refined_masks"].numpy() - z["out/refined_masks"]) * same
    return d8.max(), d8.mean(), same.mean(), dr.max()


if __name__ == "__main__":
    cases = [c for c in G.CASES if c.startswith("eval")]
    configs = [
        ("all fp32 (input fp32)", dict(input_q="none")),
        ("input fp16 only", dict()),
        ("conv x,w,out fp16", dict(conv=("fp16", "fp16", "fp16"))),
        ("conv x,w,out fp16 + rows fp16", dict(conv=("fp16", "fp16", "fp16"), rows="fp16")),
        ("conv w fp16 only", dict(conv=("none", "fp16", "none"))),
        ("conv x/out fp16 only", dict(conv=("fp16", "none", "fp16"))),
        ("conv split x,w,out", dict(conv=("split", "split", "split"), input_q="split")),
        ("conv split + rows fp16", dict(conv=("split", "split", "split"), input_q="split", rows="fp16")),
        ("conv split + rows fp16 + sparse fp16", dict(conv=("split", "split", "split"), input_q="split", rows="fp16", sparse="fp16")),
        ("conv split x,out; w fp16", dict(conv=("split", "fp16", "split"), input_q="split")),
    ]
    for name, kw in configs:
        for case in cases:
            mx, mean, agree, dr = run(case, **kw)
            print(f"{name:40s} {case:26s} a8 max {mx:.2e} mean {mean:.2e}  mask agree {agree:.5f}  refined(agree) max {dr:.2e}", flush=True)
