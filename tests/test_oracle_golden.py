"""The oracle restatement against the committed golden vectors (made from the unmodified reference by
oracle/make_golden.py). Runs without /root/reference."""
import numpy as np
import pytest
import torch

from oracle import make_golden as G
from oracle import maggie_oracle as O
from oracle import synth


def _template_shapes():
    # name -> shape is recoverable from the product model, but the oracle tests must not depend on the product:
    # rebuild the template from the golden file list + the reference-free shape table in tests/golden/shapes.npz
    z = np.load(G.GOLDEN_DIR + "/state_shapes.npz")
    return {k: torch.zeros(tuple(z[k]), dtype=torch.long if k.endswith("num_batches_tracked") else torch.float32)
            for k in z.files}


def _run_video(case):
    kw, training = G.CASES[case]
    z = np.load(G.GOLDEN_DIR + "/state_shapes_video.npz")
    P = synth.synth_state_dict({k: torch.zeros(tuple(z[k]), dtype=torch.long if k.endswith("num_batches_tracked") else torch.float32)
                                for k in z.files})
    if training:
        for k in P:
            if P[k].is_floating_point() and not k.endswith(("weight_u", "weight_v", "running_mean", "running_var")) \
                    and "dummy_downscale" not in k:
                P[k].requires_grad_(True)
    G.seed_all()
    batch = synth.make_batch(**kw)
    if training:
        out, loss = O.forward_video(P, batch, True, synth.video_cfg(), p_drop=0.0)
        loss["total"].backward()
        return out, loss, P
    with torch.no_grad():
        return O.forward_video(P, batch, False, synth.video_cfg()), None, P


@pytest.mark.parametrize("case", list(G.VIDEO_CASES))
def test_oracle_video_matches_reference_golden(case, golden):
    z = golden(case)
    out, loss, P = _run_video(case)
    for k, v in out.items():
        assert np.abs(v.detach().float().numpy() - z["out/" + k]).max() < 5e-5, k
    if loss is not None:
        for k, v in loss.items():
            assert abs(float(v) - float(z["loss/" + k])) < 1e-4 * max(1.0, abs(float(z["loss/" + k]))), k
        n = 0
        for k in z:
            if k.startswith("gradnorm/"):
                ref = float(z[k])
                assert abs(float(P[k[9:]].grad.double().norm()) - ref) < 2e-3 * ref + 1e-7, k
                n += 1
        assert n >= 300


def _run(case):
    kw, training = G.CASES[case]
    P = synth.synth_state_dict(_template_shapes())
    trainable = [k for k in P if P[k].is_floating_point() and not k.endswith(("weight_u", "weight_v", "running_mean", "running_var"))
                 and "dummy_downscale" not in k]
    if training:
        for k in trainable:
            P[k].requires_grad_(True)
    G.seed_all()
    batch = synth.make_batch(**kw)
    if training:
        out, loss, stages = O.forward(P, batch, True, synth.model_cfg(), return_stages=True,
                                      p_drop=0.0 if case in G.NO_DROPOUT else 0.1)
        loss["total"].backward()
        return out, loss, stages, P
    with torch.no_grad():
        out, stages = O.forward(P, batch, False, synth.model_cfg(), return_stages=True)
    return out, None, stages, P


@pytest.mark.parametrize("case", [c for c in G.CASES if c.startswith("eval")])  # image model
def test_oracle_eval_matches_reference_golden(case, golden):
    z = golden(case)
    out, _, stages, P = _run(case)
    for k, v in out.items():
        ref = z["out/" + k]
        assert v.shape == ref.shape
        if k == "detail_mask":
            assert (v.numpy() == ref).all()          # integer mask: bit exact
        else:
            assert np.abs(v.numpy() - ref).max() < 5e-5, k
    for k in ("os8_logits", "os8_feat", "queries", "aspp"):
        assert np.abs(stages[k].numpy() - z["stage/" + k]).max() < 2e-4, k
    for k in z:
        if k.startswith("state/"):
            assert np.abs(P[k[6:]].detach().numpy() - z[k]).max() < 1e-6


@pytest.mark.parametrize("case", [c for c in G.CASES if c.startswith("train")])
def test_oracle_train_matches_reference_golden(case, golden):
    z = golden(case)
    out, loss, _, P = _run(case)
    for k, v in out.items():
        assert np.abs(v.detach().float().numpy() - z["out/" + k]).max() < 5e-5, k
    for k, v in loss.items():
        assert abs(float(v) - float(z["loss/" + k])) < 1e-4 * max(1.0, abs(float(z["loss/" + k]))), k
    checked = 0
    for k in z:
        if k.startswith("gradnorm/"):
            g = P[k[9:]].grad
            assert g is not None, k
            ref = float(z[k])
            assert abs(float(g.double().norm()) - ref) < 2e-3 * ref + 1e-7, k
            checked += 1
    assert checked >= 290
    # exactly the four dummy_downscale weights get no gradient
    assert not any("dummy_downscale" in k for k in z if k.startswith("gradnorm/"))
