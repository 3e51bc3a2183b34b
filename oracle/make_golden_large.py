"""Goldens at the BASELINE.json config sizes and a 20-step training curve, from the UNMODIFIED reference
(/root/reference through oracle/ref_shims.py).  TEST INFRASTRUCTURE; run in the build container only:

    python -m oracle.make_golden_large            # writes tests/golden/{c2_*,c5_*,curve_*}.npz

The full-size outputs are stored compactly (the GPU box reads only these files): the OS8 logits / queries in full
(small), alpha_os8 sub-sampled (every 4th / 8th pixel), the detail mask bit-packed, losses and gradient norms as scalars.
"""
import os
import sys
import time

import numpy as np
import torch

from . import make_golden as G
from . import ref_shims, synth

LARGE = {
    # BASELINE config C2: 8 x 512 x 512 x 3 instances - evaluation forward and one training step (iter = 1)
    "c2_eval_8x512_3inst": (dict(b=8, n_f=1, n_i=3, H=512, W=512, edge_px=6.0, seed=2001), False, 4),
    "c2_train_8x512_3inst_iter1": (dict(b=8, n_f=1, n_i=3, H=512, W=512, edge_px=6.0, seed=2002, train=True, it=1), True, 4),
    # BASELINE config C5: 1024 x 1024 x 8 instances (evaluation)
    "c5_eval_1024_8inst": (dict(b=1, n_f=1, n_i=8, H=1024, W=1024, edge_px=8.0, seed=2003), False, 8),
    # BASELINE config C4 (video model MaGGIe_Temp): the 3-frame evaluation window and a 5-frame training clip at 480 x 832
    "c4_video_eval_3x480x832_2inst": (dict(b=1, n_f=3, n_i=2, H=480, W=832, edge_px=6.0, seed=2004), False, 4),
    # (the reference's LapLoss.upsample mixes up H and W - loss.py:138 - so its training runs on square crops only: the training
    #  clip is 5 x 384 x 384, about the pixel count of 480 x 832 / 2.7)
    "c4_video_train_5x384_2inst_iter1": (dict(b=1, n_f=5, n_i=2, H=384, W=384, edge_px=6.0, seed=2005, train=True, it=1), True, 4),
}
VIDEO = {"c4_video_eval_3x480x832_2inst", "c4_video_train_5x384_2inst_iter1"}
CURVE = "curve_b8_128_2inst_20steps"
# the published recipe's update rule (configs/maggie_image.yaml: AdamW lr 1.5e-4, betas (0.9, 0.999); engine/train.py:274 clips
# the global gradient norm at 0.01).  (A first version used lr 1e-3 / clip 0.1: at that step size the 8-sample BatchNorm
# dynamics are chaotic and an fp16 run separates from an fp32 one by tens of per cent within ten steps.)
CURVE_STEPS, CURVE_LR, CURVE_CLIP, CURVE_BATCHES = 20, 1.5e-4, 0.01, 4


def _reference(training, video=False):
    net = ref_shims.import_reference_network()
    model, _ = net.build_model(ref_shims.CfgNode(synth.video_cfg() if video else synth.model_cfg()))
    model.load_state_dict(synth.synth_state_dict(model.state_dict()), strict=True)
    model.train(training)
    model.decoder.inst_spec_layer.dropout.p = 0.0
    return model


def run_large(case):
    kw, training, sub = LARGE[case]
    model = _reference(training, case in VIDEO)
    batch = synth.make_batch(**kw)
    stages = {}
    model.decoder.refine_OS8.register_forward_hook(
        lambda _m, _i, out: stages.update(os8_logits=out[0].detach(), queries=out[2].detach()))
    G.seed_all()
    z = {}
    t0 = time.time()
    if training:
        out, loss = model(batch, mem_feat=None)
        loss["total"].backward()
        for k, v in loss.items():
            z["loss/" + k] = np.float64(float(v))
        for n, p in model.named_parameters():
            if p.grad is not None:
                z["gradnorm/" + n] = np.float64(float(p.grad.double().norm()))
    else:
        with torch.no_grad():
            out = model(batch, mem_feat=None)
    z["stage/os8_logits"] = stages["os8_logits"].float().numpy()
    z["stage/queries"] = stages["queries"].float().numpy()
    z["sub"] = np.int64(sub)
    for k in ("alpha_os8", "refined_masks", "temp_alpha", "diff_pred_forward", "diff_pred_backward"):
        if k in out:
            z["out_sub/" + k] = out[k].detach().float().numpy()[..., ::sub, ::sub].copy()
    if "mem_feat" in out and torch.is_tensor(out["mem_feat"]):
        z["out_sub/mem_feat"] = out["mem_feat"].detach().float().numpy()[..., ::4, ::4].copy()
    dm = out["detail_mask"].detach().numpy() != 0
    z["out_bits/detail_mask"] = np.packbits(dm.reshape(-1))
    z["out_shape/detail_mask"] = np.array(dm.shape, np.int64)
    z["detail_count"] = np.int64(dm.sum())
    print(f"{case}: reference ran in {time.time() - t0:.1f} s, active fraction {dm.mean():.4f}")
    return z


def curve_batches():
    return [synth.make_batch(b=8, n_f=1, n_i=2, H=128, W=128, edge_px=4.0, seed=3000 + i, train=True, it=100000)
            for i in range(CURVE_BATCHES)]


def run_curve():
    """20 optimizer steps of the reference engine's update rule (engine/train.py:265-283 without the fp16 scaler:
    backward, clip_grad_norm_(all parameters), AdamW step) on a cycle of 4 batches; iter = 100000 (predicted-alpha regime)."""
    model = _reference(True)
    opt = torch.optim.AdamW(model.parameters(), lr=CURVE_LR, betas=(0.9, 0.999), weight_decay=0.01)
    batches = curve_batches()
    rows, norms = [], []
    keys = None
    for step in range(CURVE_STEPS):
        G.seed_all(G.RNG_SEED + step)
        opt.zero_grad(set_to_none=True)
        _, loss = model(batches[step % CURVE_BATCHES], mem_feat=None)
        loss["total"].backward()
        gn = torch.nn.utils.clip_grad_norm_(model.parameters(), CURVE_CLIP)
        opt.step()
        keys = keys or sorted(loss)
        rows.append([float(loss[k]) for k in keys])
        norms.append(float(gn))
        print(f"  step {step:2d}: total {float(loss['total']):.5f}  grad norm {float(gn):.4f}", flush=True)
    return {"keys": np.array(keys), "losses": np.array(rows, np.float64), "grad_norm": np.array(norms, np.float64),
            "lr": np.float64(CURVE_LR), "clip": np.float64(CURVE_CLIP)}


def main():
    only = sys.argv[1:]
    for case in LARGE:
        if only and case not in only:
            continue
        z = run_large(case)
        path = os.path.join(G.GOLDEN_DIR, case + ".npz")
        np.savez_compressed(path, **z)
        print(f"  -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")
    if not only or CURVE in only:
        z = run_curve()
        path = os.path.join(G.GOLDEN_DIR, CURVE + ".npz")
        np.savez_compressed(path, **z)
        print(f"  -> {path}")


if __name__ == "__main__":
    sys.exit(main())
