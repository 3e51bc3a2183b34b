"""-m gpu: parity at the BASELINE.json config sizes against goldens of the UNMODIFIED reference (oracle/make_golden_large.py,
stored compactly), and a 20-step training curve against the reference engine's update rule.

  C2 (8 x 512 x 512 x 3 instances): evaluation forward in both precisions, one training step (iter = 1);
  C5 (1024 x 1024 x 8 instances): evaluation forward;
  C4 (video model): the 3-frame evaluation window at 480 x 832 x 2 instances, a 5-frame 384 x 384 training clip (the reference's
      Laplacian loss only runs on square crops);
  loss curve: 20 steps (backward, clip_grad_norm_, AdamW) on a cycle of four 8 x 128 x 128 batches, iter = 100000.
Tolerances are stated next to each assert; measured values are printed with `-s`."""
import os

import numpy as np
import pytest
import torch

from maggie_b200.config import CfgNode
from maggie_b200.network import build_model
from oracle import make_golden as G
from oracle import make_golden_large as GL
from oracle import synth

pytestmark = pytest.mark.gpu
ALPHA_TOL = 1e-3          # north star: alpha within 1e-3 max abs (met by precision="high", the reference evaluates in fp32)


def _model(training, precision="fp16"):
    m, _ = build_model(CfgNode(synth.model_cfg()))
    m.load_state_dict(synth.synth_state_dict(m.state_dict()), strict=True)
    m.decoder.inst_spec_layer.dropout.p = 0.0
    return m.cuda().train(training).set_precision(precision)


def _dev(batch):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}


def _detail(z):
    shape = tuple(int(v) for v in z["out_shape/detail_mask"])
    return np.unpackbits(z["out_bits/detail_mask"])[:int(np.prod(shape))].reshape(shape).astype(bool)


@pytest.mark.parametrize("case", ["c2_eval_8x512_3inst", "c5_eval_1024_8inst"])
def test_eval_at_baseline_config_size(case, golden):
    kw, _, sub = GL.LARGE[case]
    z = golden(case)
    res = {}
    for precision in ("high", "fp16"):
        m = _model(False, precision)
        grabbed = {}
        imd = m.decoder.refine_OS8.forward
        m.decoder.refine_OS8.forward = lambda *a, **k: grabbed.setdefault("out", imd(*a, **k))
        G.seed_all()
        with torch.no_grad():
            out = m(_dev(synth.make_batch(**kw)), mem_feat=None)
        a8 = out["alpha_os8"].float().cpu().numpy()[..., ::sub, ::sub]
        lg = grabbed["out"][0].float().cpu().numpy()
        res[precision] = (float(np.abs(a8 - z["out_sub/alpha_os8"]).max()), float(np.abs(lg - z["stage/os8_logits"]).max()),
                          float((out["detail_mask"].cpu().numpy().astype(bool) == _detail(z)).mean()))
        del m, out
        torch.cuda.empty_cache()
    print(f"\n{case}: max|alpha_os8 - ref| (every {sub}th pixel), max|os8 logits - ref|, detail-mask agreement")
    for k, v in res.items():
        print(f"  {k:5s} {v[0]:.3e}  {v[1]:.3e}  {v[2]:.6f}")
    assert res["high"][0] <= ALPHA_TOL, res
    assert res["high"][2] >= 0.9995
    assert res["fp16"][0] <= 2.5e-2          # fp16 storage floor at this size (measured ~1e-2; noise realisation dependent)
    assert res["fp16"][0] > 5 * res["high"][0]


def test_c2_training_step_matches_the_reference(golden):
    """BASELINE config C2, one training step at iter = 1: losses, active sites, gradient norms (fp16 autocast width)."""
    case = "c2_train_8x512_3inst_iter1"
    kw, _, sub = GL.LARGE[case]
    z = golden(case)
    m = _model(True)
    G.seed_all()
    out, loss = m(_dev(synth.make_batch(**kw)), mem_feat=None)
    (loss["total"] * 64.0).backward()
    assert int(out["detail_mask"].sum()) == int(z["detail_count"])      # the uncertain region comes from the GT: bit exact
    assert bool((out["detail_mask"].cpu().numpy().astype(bool) == _detail(z)).all())
    print()
    for k in sorted(k[5:] for k in z if k.startswith("loss/")):
        ref, got = float(z["loss/" + k]), float(loss[k])
        print(f"  {k:16s} ref {ref:.5f} ours {got:.5f}  rel {abs(got - ref) / max(abs(ref), 1e-9):.2e}")
        assert abs(got - ref) <= 2e-3 * max(1.0, abs(ref)), (k, got, ref)   # measured <= 5e-4 (fp16 activations vs fp32 reference)
    # (training-mode alphas are ill-conditioned: batch statistics over 8 samples amplify the fp16 rounding of the input;
    #  the losses above, which integrate over all pixels, agree to 5e-4 - measured mean |d alpha| 8.4e-3)
    a8 = out["alpha_os8"].detach().float().cpu().numpy()[..., ::sub, ::sub]
    assert float(np.abs(a8 - z["out_sub/alpha_os8"]).mean()) < 1.25 * 8.5e-3
    rel = []
    for n, p in m.named_parameters():
        key = "gradnorm/" + n
        if key in z and p.grad is not None and float(z[key]) > 1e-4:
            rel.append(abs(float((p.grad.double() / 64.0).norm()) - float(z[key])) / float(z[key]))
    rel = np.array(rel)
    print(f"  gradient norms of {len(rel)} parameters: median rel err {np.median(rel):.3e}, 90th pct {np.percentile(rel, 90):.3e}")
    assert len(rel) > 250 and np.median(rel) < 0.05 and np.percentile(rel, 90) < 0.25


def test_loss_curve_tracks_the_reference(golden):
    """20 optimizer steps with the fused optimizer tail (K14) on the flat gradient buffer against the reference model under
    torch's clip_grad_norm_ + AdamW (fp32, CPU): the loss curve must track within the stated tolerance."""
    from maggie_b200.dp import FlatGradAllReduce
    from maggie_b200.optim import FusedAdamW
    z = golden(GL.CURVE)
    keys = [str(k) for k in z["keys"]]
    ref = z["losses"][:, keys.index("total")]
    m = _model(True)
    flat = FlatGradAllReduce(m.parameters())
    opt = FusedAdamW(flat, lr=float(z["lr"]), betas=(0.9, 0.999), weight_decay=0.01, clip_norm=float(z["clip"]))
    batches = [_dev(b) for b in GL.curve_batches()]
    scale, ours, norms = 256.0, [], []
    for step in range(GL.CURVE_STEPS):
        G.seed_all(G.RNG_SEED + step)
        flat.zero()
        _, loss = m(batches[step % GL.CURVE_BATCHES], mem_feat=None)
        (loss["total"] * scale).backward()
        rep = opt.step(grad_scale=scale)
        ours.append(float(loss["total"]))
        norms.append(float(rep[0]))
    ours, norms = np.array(ours), np.array(norms)
    rel = np.abs(ours - ref) / ref
    print("\n  step   reference   ours      rel.err   |grad| ref / ours")
    for i in range(GL.CURVE_STEPS):
        print(f"  {i:3d}   {ref[i]:9.5f}  {ours[i]:9.5f}  {rel[i]:.3e}   {z['grad_norm'][i]:.3f} / {norms[i]:.3f}")
    # stated tolerance: every step within 6 % of the reference curve and 1.2 % on average, and the same overall descent
    # (9.5 -> 5.0).  Measured on B200 over 12 runs (fp16 activations against the fp32 reference; the training backward is
    # not bit-reproducible - fp32 atomics in the BatchNorm / weight-gradient reductions - and 20 optimizer steps amplify
    # the last-bit differences): per-run max 1.6 % .. 3.8 % (always at one of the last steps), mean 0.4 % .. 0.7 %
    assert rel.max() < 0.06 and rel.mean() < 0.012, (rel.max(), rel.mean())
    assert abs((ours[-4:].mean() / ours[:4].mean()) - (ref[-4:].mean() / ref[:4].mean())) < 0.02
    assert np.abs(norms - z["grad_norm"]).max() < 0.25 * z["grad_norm"].max()


# ---- video model (MaGGIe_Temp) at the BASELINE C4 size -----------------------------------------------------------------
def _video_model(training):
    m, _ = build_model(CfgNode(synth.video_cfg()))
    m.load_state_dict(synth.synth_state_dict(m.state_dict()), strict=True)
    m.decoder.inst_spec_layer.dropout.p = 0.0
    return m.cuda().train(training)


def test_video_eval_window_at_c4_size(golden):
    """BASELINE config C4, the 3-frame evaluation window at 480 x 832 x 2 instances against the unmodified reference (fp16 storage
    against the fp32 reference: the bounds are the 16-bit floor of this model, measured values printed with -s)."""
    case = "c4_video_eval_3x480x832_2inst"
    kw, _, sub = GL.LARGE[case]
    z = golden(case)
    m = _video_model(False)
    grabbed = {}
    imd = m.decoder.refine_OS8.forward
    m.decoder.refine_OS8.forward = lambda *a, **k: grabbed.setdefault("out", imd(*a, **k))
    G.seed_all()
    with torch.no_grad():
        out = m(_dev(synth.make_batch(**kw)), mem_feat=None)
    s = lambda t: t.float().cpu().numpy()[..., ::sub, ::sub]
    err = {k: np.abs(s(out[k]) - z["out_sub/" + k]) for k in ("alpha_os8", "refined_masks", "temp_alpha", "diff_pred_forward",
                                                              "diff_pred_backward")}
    lg = np.abs(grabbed["out"][0].float().cpu().numpy() - z["stage/os8_logits"])
    mem = np.abs(out["mem_feat"].float().cpu().numpy()[..., ::4, ::4] - z["out_sub/mem_feat"])
    agree = float((out["detail_mask"].cpu().numpy().astype(bool) == _detail(z)).mean())
    print(f"\n{case}: detail-mask agreement {agree:.5f}; |os8 logits| err max {lg.max():.3e}; mem_feat err max {mem.max():.3e} mean {mem.mean():.3e}")
    for k, e in err.items():
        print(f"  {k:20s} max {e.max():.3e}  mean {e.mean():.3e}  frac > 1e-2: {np.mean(e > 1e-2):.4f}")
    # measured on B200 (the evaluation forward is bit-reproducible): agreement 1.0; alpha_os8 max 3.4e-3 / mean 1.9e-4; refined
    # alpha max 3.7e-2 (next to the refinement band's edge) / mean 7.6e-5; temp_alpha max 4.8e-4; difference maps max 4.8e-4;
    # GRU hidden state max 2.0e-3 / mean 2.5e-4.  Bounds = 1.5 x measured.
    assert agree > 0.9999
    assert err["alpha_os8"].max() < 5.1e-3 and err["alpha_os8"].mean() < 2.8e-4
    assert err["refined_masks"].max() < 5.5e-2 and err["refined_masks"].mean() < 1.2e-4
    assert err["temp_alpha"].max() < 7.2e-4 and err["temp_alpha"].mean() < 1.1e-4
    assert err["diff_pred_forward"].max() < 7.2e-4 and err["diff_pred_backward"].max() < 5.2e-4
    assert mem.max() < 3e-3 and mem.mean() < 3.8e-4


def test_video_training_clip_matches_the_reference(golden):
    """A 5-frame 384 x 384 x 2 training clip (iter = 1) of the video model: losses (incl. the temporal ones), active sites and
    gradient norms against the unmodified reference."""
    case = "c4_video_train_5x384_2inst_iter1"
    kw, _, sub = GL.LARGE[case]
    z = golden(case)
    m = _video_model(True)
    G.seed_all()
    out, loss = m(_dev(synth.make_batch(**kw)), mem_feat=None)
    (loss["total"] * 64.0).backward()
    assert int(out["detail_mask"].sum()) == int(z["detail_count"])
    print()
    worst = 0.0
    for k in sorted(k[5:] for k in z if k.startswith("loss/")):
        ref, got = float(z["loss/" + k]), float(loss[k])
        rel = abs(got - ref) / max(abs(ref), 1e-3)
        worst = max(worst, rel)
        print(f"  {k:18s} ref {ref:.5f} ours {got:.5f}  rel {rel:.2e}")
    assert worst < 5e-3, worst                    # measured 1.3e-3 (loss_dtSSD_os8); fp16 activations against the fp32 reference
    rel = []
    for n, p in m.named_parameters():
        key = "gradnorm/" + n
        if key in z and p.grad is not None and float(z[key]) > 1e-4:
            rel.append(abs(float((p.grad.double() / 64.0).norm()) - float(z[key])) / float(z[key]))
    rel = np.array(rel)
    print(f"  gradient norms of {len(rel)} parameters: median rel err {np.median(rel):.3e}, 90th pct {np.percentile(rel, 90):.3e}")
    assert len(rel) > 250 and np.median(rel) < 0.015 and np.percentile(rel, 90) < 0.1      # measured 3.1e-3 / 2.5e-2
