"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference through
oracle/ref_shims.py) on the synthetic cases below, and report oracle-vs-reference differences.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container only:
    python -m oracle.make_golden            # writes tests/golden/<case>.npz
The GPU box has no /root/reference; tests there read the committed .npz files.
"""
import os
import random
import sys

import numpy as np
import torch

from . import maggie_oracle as O
from . import ref_shims, synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (make_batch kwargs, training)
CASES = {
    "eval_c1_256_1inst": (dict(b=1, n_f=1, n_i=1, H=256, W=256, edge_px=6.0, seed=1234), False),
    "eval_192x256_3inst": (dict(b=1, n_f=1, n_i=3, H=192, W=256, edge_px=4.0, seed=4321), False),
    "eval_128_3inst_maskos8": (dict(b=2, n_f=1, n_i=3, H=128, W=128, edge_px=3.0, seed=99, mask_os8=True), False),
    "train_128_2inst_iter1": (dict(b=2, n_f=1, n_i=2, H=128, W=128, edge_px=4.0, seed=777, train=True, it=1), True),
    "train_128_3inst_iter100k": (dict(b=2, n_f=1, n_i=3, H=128, W=128, edge_px=4.0, seed=778, train=True, it=100000), True),
    # better-conditioned training case for the 16-bit GPU path: 8 samples per BatchNorm batch (the ASPP global-pool BN
    # sees only `b` values per channel), inst_spec dropout off (CUDA and CPU dropout streams differ)
    "train_b8_128_2inst_nodrop": (dict(b=8, n_f=1, n_i=2, H=128, W=128, edge_px=4.0, seed=779, train=True, it=100000), True),
    # degenerate training batch: binary alphas -> empty uncertain set -> the reference paints its dummy 50x50 patch into
    # every one of the 10 slots (resnet_inst_matt_spconv.py:347-348)
    "train_256_2inst_empty_roi": (dict(b=2, n_f=1, n_i=2, H=256, W=256, edge_px=4.0, seed=780, train=True, it=1, binary_alpha=True), True),
}
# video model (MaGGIe_Temp): 3-frame eval window (with and without prev_pred) and a training clip
VIDEO_CASES = {
    "video_eval_3f_128x192_2inst": (dict(b=1, n_f=3, n_i=2, H=128, W=192, edge_px=4.0, seed=501), False),
    "video_train_4f_128_2inst_nodrop": (dict(b=2, n_f=4, n_i=2, H=128, W=128, edge_px=4.0, seed=502, train=True, it=100000), True),
}
CASES.update(VIDEO_CASES)
NO_DROPOUT = {"train_b8_128_2inst_nodrop", "video_train_4f_128_2inst_nodrop", "train_256_2inst_empty_roi"}
RNG_SEED = 2024
SMALL_GRAD_NUMEL = 2048


def seed_all(seed=RNG_SEED):
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)


def cfg_for(case):
    return synth.video_cfg() if case in VIDEO_CASES else synth.model_cfg()


def build_reference(training, case=None):
    net = ref_shims.import_reference_network()
    model, _ = net.build_model(ref_shims.CfgNode(cfg_for(case)))
    model.load_state_dict(synth.synth_state_dict(model.state_dict()), strict=True)
    model.train(training)
    return model


def run_reference(case):
    kw, training = CASES[case]
    model = build_reference(training, case)
    if case in NO_DROPOUT:
        model.decoder.inst_spec_layer.dropout.p = 0.0
    batch = synth.make_batch(**kw)
    stages = {}

    def hook_imd(_m, _i, out):
        stages["os8_logits"], stages["os8_feat"], stages["queries"] = out[0].detach(), out[1].detach(), out[2].detach()

    def hook_aspp(_m, _i, out):
        stages["aspp"] = out.detach()

    model.decoder.refine_OS8.register_forward_hook(hook_imd)
    model.aspp.register_forward_hook(hook_aspp)
    seed_all()
    if training:
        out, loss = model(batch, mem_feat=None)
        loss["total"].backward()
        grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
        return out, loss, stages, grads, model
    with torch.no_grad():
        out = model(batch, mem_feat=None)
    return out, None, stages, None, model


def run_oracle(case):
    kw, training = CASES[case]
    net = ref_shims.import_reference_network()
    tmpl, _ = net.build_model(ref_shims.CfgNode(cfg_for(case)))
    P = synth.synth_state_dict(tmpl.state_dict())
    names = {n for n, p in tmpl.named_parameters() if p.requires_grad}
    if case in VIDEO_CASES:
        return run_oracle_video(case, P, names)
    for n in P:
        if training and n in names and P[n].is_floating_point():
            P[n].requires_grad_(True)
    batch = synth.make_batch(**kw)
    seed_all()
    if training:
        out, loss, stages = O.forward(P, batch, True, synth.model_cfg(), return_stages=True,
                                      p_drop=0.0 if case in NO_DROPOUT else 0.1)
        loss["total"].backward()
        grads = {n: p.grad for n, p in P.items() if p.requires_grad and p.grad is not None}
        return out, loss, stages, grads, P
    with torch.no_grad():
        out, stages = O.forward(P, batch, False, synth.model_cfg(), return_stages=True)
    return out, None, stages, None, P


def run_oracle_video(case, P, names):
    kw, training = CASES[case]
    for n in P:
        if training and n in names and P[n].is_floating_point():
            P[n].requires_grad_(True)
    batch = synth.make_batch(**kw)
    seed_all()
    stages = {k: torch.zeros(1) for k in ("os8_logits", "os8_feat", "queries", "aspp")}
    if training:
        out, loss = O.forward_video(P, batch, True, synth.video_cfg(), p_drop=0.0)
        loss["total"].backward()
        grads = {n: p.grad for n, p in P.items() if p.requires_grad and p.grad is not None}
        return out, loss, stages, grads, P
    with torch.no_grad():
        out = O.forward_video(P, batch, False, synth.video_cfg())
    return out, None, stages, None, P


def pack(out, loss, stages, grads, state):
    z = {}
    for k, v in out.items():
        z["out/" + k] = v.detach().float().numpy()
    for k in ("os8_logits", "os8_feat", "queries", "aspp"):
        if k in stages:
            z["stage/" + k] = stages[k].detach().float().numpy()
    if loss is not None:
        for k, v in loss.items():
            z["loss/" + k] = np.float64(float(v))
        for n, g in grads.items():
            z["gradnorm/" + n] = np.float64(float(g.double().norm()))
            if g.numel() <= SMALL_GRAD_NUMEL:
                z["grad/" + n] = g.float().numpy()
    # post-forward SpectralNorm state for two layers (u,v are mutated by every forward)
    for n in ("encoder.conv1.module.weight_u", "decoder.layer2.0.conv1.module.weight_v"):
        z["state/" + n] = state[n].detach().float().numpy()
    return z


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    only = sys.argv[1:]
    for case in CASES:
        if only and case not in only:
            continue
        ref = run_reference(case)
        z = pack(*ref[:4], ref[4].state_dict())
        np.savez_compressed(os.path.join(GOLDEN_DIR, case + ".npz"), **z)
        orc = run_oracle(case)
        zo = pack(*orc[:4], orc[4])
        worst = {}
        for k in z:
            if k not in zo:
                print("  oracle missing", k)
                continue
            d = float(np.max(np.abs(np.asarray(z[k], np.float64) - np.asarray(zo[k], np.float64)))) if np.size(z[k]) else 0.0
            grp = k.split("/")[0]
            worst[grp] = max(worst.get(grp, 0.0), d / (1.0 if grp != "gradnorm" else max(1e-12, abs(float(z[k])))))
        sz = os.path.getsize(os.path.join(GOLDEN_DIR, case + ".npz")) / 1e6
        print(f"{case}: {sz:.2f} MB; oracle-vs-reference max abs diff per group: "
              + ", ".join(f"{g}={v:.2e}" for g, v in worst.items()))


if __name__ == "__main__":
    sys.exit(main())
