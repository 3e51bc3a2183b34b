"""-m gpu: the whole forward/backward path on the B200 against the reference goldens (fp16 compute vs the fp32
reference).

Measured noise floor (tools/debug_parity.py, B200): with 16-bit storage of activations the OS8 alpha differs from the
fp32 reference by 6e-3 .. 1.2e-2 max-abs (2e-4 .. 5e-4 mean-abs) on these random-weight models - the same level as
torch's own cuDNN fp16 path with identical host code (7e-3 .. 1.2e-2) - while an fp32 run of the same host code on
the GPU agrees to 2e-4.  The 1e-3 max-abs target of the north star is therefore not met by ANY 16-bit path on these
weights; the bounds below are 2x the measured floor.  Training-mode forward is ill-conditioned on tiny batches
(BatchNorm over 2..8 values per channel in the ASPP global-pool branch): rounding just the 6-channel input to fp16
moves alpha_os8 by 6e-2 max / 2.4e-3 mean in exact fp32 arithmetic (reproduced on CPU), so training parity is
asserted on losses, mean errors and gradient norms, and per-op (tests/test_gpu_conv_bn.py) where it is tight."""
import numpy as np
import pytest
import torch

from maggie_b200 import _lib
from maggie_b200.config import CfgNode
from maggie_b200.network import build_model
from oracle import make_golden as G
from oracle import synth

pytestmark = pytest.mark.gpu
ALPHA_TOL = 1e-3   # north star; met by precision="high" (asserted below and in tests/test_gpu_highprec.py)
# measured max-abs alpha_os8 error of the fp16 mode on the eval goldens (B200, bit-reproducible): asserted at 1.25x
FP16_FLOOR = {"eval_c1_256_1inst": 1.05e-2, "eval_192x256_3inst": 1.05e-2, "eval_128_3inst_maskos8": 1.2e-2}


def _model(training):
    m, _ = build_model(CfgNode(synth.model_cfg()))
    m.load_state_dict(synth.synth_state_dict(m.state_dict()), strict=True)
    return m.cuda().train(training)


def _to_dev(batch):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}


@pytest.mark.parametrize("case", [c for c in G.CASES if c.startswith("eval")])
def test_eval_alpha_parity(case, golden):
    kw, _ = G.CASES[case]
    z, m = golden(case), _model(False)
    G.seed_all()
    _lib.reset_launch_count()
    with torch.no_grad():
        out = m(_to_dev(synth.make_batch(**kw)), mem_feat=None)
    out = {k: v.float().cpu().numpy() for k, v in out.items()}
    assert _lib.launch_count() > 0, "native library was not used"
    d8 = np.abs(out["alpha_os8"] - z["out/alpha_os8"])
    assert d8.max() <= 1.25 * FP16_FLOOR[case] and d8.mean() < 6e-4, f"alpha_os8 max abs diff {d8.max()}, mean {d8.mean()}"
    # the stated tolerance holds in the fp32-accurate evaluation mode (same weights, same batch)
    G.seed_all()
    with torch.no_grad():
        # (a fresh model: every forward advances the spectral-norm power iteration)
        hi = _model(False).set_precision("high")(_to_dev(synth.make_batch(**kw)), mem_feat=None)
    assert np.abs(hi["alpha_os8"].float().cpu().numpy() - z["out/alpha_os8"]).max() <= ALPHA_TOL
    same = out["detail_mask"] == z["out/detail_mask"]
    assert same.mean() > 0.995, f"detail masks agree on {same.mean()}"
    for k in ("alpha_os4", "alpha_os1", "refined_masks"):
        # compare away from pixels whose own or whose neighbours' mask membership flipped
        d = np.abs(out[k] - z["out/" + k])
        frac_bad = (d > 1e-2).mean()
        assert frac_bad < 0.02 and d.mean() < 3e-3, f"{k}: {frac_bad:.4f} of pixels differ by more than 1e-2 (mean {d.mean():.3e})"


@pytest.mark.parametrize("case", [c for c in G.CASES if c.startswith("train")])
def test_train_loss_and_gradient_parity(case, golden):
    kw, _ = G.CASES[case]
    z, m = golden(case), _model(True)
    m.decoder.inst_spec_layer.dropout.p = 0.0  # CUDA and CPU dropout streams differ; goldens used p=0.1 -> loose tol below
    G.seed_all()
    out, loss = m(_to_dev(synth.make_batch(**kw)), mem_feat=None)
    scale = 64.0
    (loss["total"] * scale).backward()
    for k in ("loss_rec_os8", "loss_lap_os8", "loss_grad_os8", "loss_max_atten"):
        ref = float(z["loss/" + k])
        assert abs(float(loss[k]) - ref) < 0.03 * max(1.0, abs(ref)), (k, float(loss[k]), ref)
    d8 = np.abs(out["alpha_os8"].detach().float().cpu().numpy() - z["out/alpha_os8"]).mean()
    assert d8 < 2.5e-2, d8
    # encoder / dense-decoder gradients do not depend on the dropout mask only through the sparse branch
    checked = 0
    for k, p in m.named_parameters():
        if "gradnorm/" + k in z and k.startswith(("decoder.refine_OS8.token", "decoder.refine_OS8.final", "decoder.refine_OS8.query")):
            ref = float(z["gradnorm/" + k])
            got = float((p.grad.double() / scale).norm())
            assert abs(got - ref) < 0.2 * ref + 1e-5, (k, got, ref)
            checked += 1
    assert checked >= 10
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)


def test_empty_mask_raises_like_the_reference():
    """A sample without any mask pixel: the reference raises ValueError("Mask is empty") from its NaN checks
    (module/mask_attention.py:95-98); here a device-side flag travels with the step's one host read."""
    for training in (False, True):
        m = _model(training)
        batch = _to_dev(synth.make_batch(b=2, n_f=1, n_i=2, H=128, W=128, edge_px=4.0, seed=8, train=training, it=1))
        batch["mask"][1] = 0
        G.seed_all()
        with pytest.raises(ValueError, match="Mask is empty"):
            with torch.set_grad_enabled(training):
                m(batch, mem_feat=None)


def test_post_warmup_step_reads_the_host_once():
    """iter >= 3 * warmup_detail_iter: the predicted OS8 alpha guides the detail stage (resnet_inst_matt_spconv.py:311-316).
    The reference tests `x_os8.sum() == 0` on the host; here the switch is a device flag and the step's only synchronising
    call is the 32-byte read of the status word (site counts + flags)."""
    import warnings
    m = _model(True)
    m.decoder.inst_spec_layer.dropout.p = 0.0
    batch = _to_dev(synth.make_batch(b=2, n_f=1, n_i=2, H=128, W=128, edge_px=4.0, seed=3, train=True, it=100000))
    for _ in range(2):      # caches (index tensors, pinned staging) are filled by the first steps
        G.seed_all()
        _, loss = m(batch, mem_feat=None)
        (loss["total"] * 64.0).backward()
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("warn")
    try:
        with warnings.catch_warnings(record=True) as rec:
            warnings.simplefilter("always")
            G.seed_all()
            out, loss = m(batch, mem_feat=None)
            (loss["total"] * 64.0).backward()
    finally:
        torch.cuda.set_sync_debug_mode("default")
    syncs = [w for w in rec if "synchroniz" in str(w.message).lower()]
    assert len(syncs) == 1, [str(w.message) for w in syncs]
    assert np.isfinite(float(loss["total"]))
    # an all-zero prediction switches to the ground truth on the device (and the host learns it from the same read)
    use_gt, unk, T = m.decoder._guidance_and_roi(torch.zeros_like(batch["alpha"][:, 0]), batch["alpha"][:, 0].float(), 100000, None,
                                                 __import__("maggie_b200").ops.new_status(batch["alpha"].device))
    assert use_gt and T.counts[0] > 0


def test_full_size_c2_properties():
    """BASELINE config C2 (8 x 512 x 512 x 3 instances, training): size-independent properties."""
    m = _model(True)
    batch = _to_dev(synth.make_batch(b=8, n_f=1, n_i=3, H=512, W=512, edge_px=6.0, train=True, it=1))
    G.seed_all()
    out, loss = m(batch, mem_feat=None)
    (loss["total"] * 64.0).backward()
    for k, v in out.items():
        assert v.shape == (8, 1, 3, 512, 512), k
        assert bool(torch.isfinite(v.float()).all()), k
    for k in ("alpha_os8", "alpha_os4", "alpha_os1", "refined_masks"):
        assert float(out[k].min()) >= 0.0 and float(out[k].max()) <= 1.0
    dm = out["detail_mask"].bool()
    # outside the refined region the fused alpha IS the OS8 alpha
    assert bool((out["refined_masks"][~dm] == out["alpha_os8"][~dm]).all())
    n1 = m.last_site_counts[0]
    assert n1 == int(dm.sum())
    assert 0.02 < n1 / (8 * 3 * 512 * 512) < 0.2
    assert np.isfinite(float(loss["total"]))
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)


def test_cuda_graph_dense_stage_matches_eager():
    """The graph-replayed dense stage (forward + backward) gives the same losses and gradients as the eager run."""
    batch = _to_dev(synth.make_batch(b=2, n_f=1, n_i=2, H=128, W=128, edge_px=4.0, seed=3, train=True, it=1))
    res = []
    for graphs in (False, True):
        m = _model(True)
        m.decoder.inst_spec_layer.dropout.p = 0.0
        m.enable_cuda_graphs(graphs)
        out = None
        for _ in range(2):   # second call replays the captured graphs
            for p in m.parameters():
                p.grad = None
            G.seed_all()
            out, loss = m(batch, mem_feat=None)
            (loss["total"] * 64.0).backward()
        gn = torch.stack([p.grad.float().norm() for p in m.parameters() if p.grad is not None])
        res.append((float(loss["total"]), float(loss["loss_max_atten"]), gn, m.state_dict()["encoder.conv1.module.weight_u"].clone()))
    (l0, a0, g0, u0), (l1, a1, g1, u1) = res
    assert abs(l0 - l1) < 2e-2 * abs(l0) and abs(a0 - a1) < 2e-2 * abs(a0), (l0, l1, a0, a1)
    assert g0.shape == g1.shape
    rel = ((g0 - g1).abs() / (g0.abs() + 1e-3 * g0.abs().max()))
    assert float(rel.median()) < 2e-2, float(rel.median())


def test_eval_graph_replay_is_bit_identical_to_eager_and_keeps_the_state_in_step():
    """Evaluation forward with the dense stage replayed as a CUDA graph: identical bits to the eager run on every call,
    and the state the forward advances (spectral-norm power iteration) evolves identically - the warm-up passes of the
    capture do not count."""
    batch = _to_dev(synth.make_batch(b=1, n_f=1, n_i=2, H=128, W=128, edge_px=4.0, seed=5))
    outs, us = [], []
    for graphs in (False, True):
        m = _model(False)
        m.enable_cuda_graphs(graphs)
        with torch.no_grad():
            seq = [m(batch, mem_feat=None) for _ in range(3)]        # graphs: capture + first replay, then two replays
        outs.append([{k: v.float().clone() for k, v in o.items()} for o in seq])
        us.append(m.state_dict()["encoder.conv1.module.weight_u"].clone())
    for a, b in zip(*outs):
        for k in ("alpha_os8", "refined_masks", "detail_mask"):
            assert torch.equal(a[k], b[k]), k
    assert torch.equal(us[0], us[1])
    # the three forwards really differ (the power iteration moves the weights): the comparison above is not vacuous
    assert not torch.equal(outs[0][0]["alpha_os8"], outs[0][2]["alpha_os8"])


def test_training_graph_capture_does_not_count_its_warmup_passes():
    """BatchNorm running statistics after ONE graph-replayed training step equal those after one eager step."""
    batch = _to_dev(synth.make_batch(b=2, n_f=1, n_i=2, H=128, W=128, edge_px=4.0, seed=3, train=True, it=1))
    stats = []
    for graphs in (False, True):
        m = _model(True)
        m.enable_cuda_graphs(graphs)
        G.seed_all()
        _, loss = m(batch, mem_feat=None)
        (loss["total"] * 64.0).backward()
        sd = m.state_dict()
        stats.append((sd["encoder.bn1.running_mean"].clone(), sd["encoder.bn1.num_batches_tracked"].clone(),
                      sd["encoder.conv1.module.weight_u"].clone()))
    (rm0, n0, u0), (rm1, n1, u1) = stats
    assert int(n0) == int(n1) == 1
    assert torch.allclose(rm0, rm1, rtol=1e-3, atol=1e-5)
    assert torch.allclose(u0, u1, rtol=1e-4, atol=1e-6)


# ---- video model (MaGGIe_Temp, BASELINE config C4) -------------------------------------------------------------
def _video_model(training):
    m, _ = build_model(CfgNode(synth.video_cfg()))
    m.load_state_dict(synth.synth_state_dict(m.state_dict()), strict=True)
    m.decoder.inst_spec_layer.dropout.p = 0.0
    return m.cuda().train(training)


def test_video_eval_parity(golden):
    case = "video_eval_3f_128x192_2inst"
    kw, _ = G.CASES[case]
    z, m = golden(case), _video_model(False)
    G.seed_all()
    with torch.no_grad():
        out = m(_to_dev(synth.make_batch(**kw)), mem_feat=None)
    out = {k: v.float().cpu().numpy() for k, v in out.items()}
    d8 = np.abs(out["alpha_os8"] - z["out/alpha_os8"])
    assert d8.max() < 5e-2 and d8.mean() < 2e-3, (d8.max(), d8.mean())     # 16-bit noise floor + the >=0.95 -> 1 snap
    assert (out["detail_mask"] == z["out/detail_mask"]).mean() > 0.99
    assert np.abs(out["mem_feat"] - z["out/mem_feat"]).mean() < 5e-3
    for k in ("diff_pred_forward", "diff_pred_backward"):
        assert np.abs(out[k] - z["out/" + k]).mean() < 5e-3, k
    assert np.abs(out["refined_masks"] - z["out/refined_masks"]).mean() < 5e-3


def test_video_train_loss_parity(golden):
    case = "video_train_4f_128_2inst_nodrop"
    kw, _ = G.CASES[case]
    z, m = golden(case), _video_model(True)
    G.seed_all()
    out, loss = m(_to_dev(synth.make_batch(**kw)), mem_feat=None)
    (loss["total"] * 64.0).backward()
    for k in ("loss_rec_os8", "loss_max_atten", "loss_temp_bce", "loss_dtSSD_os8", "loss_lap_os8"):
        ref = float(z["loss/" + k])
        assert abs(float(loss[k]) - ref) < 0.05 * max(1.0, abs(ref)), (k, float(loss[k]), ref)
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
    gru = m.decoder.os8_temp_module.ih[0].weight.grad
    assert gru is not None and float(gru.abs().sum()) > 0


def test_video_c4_shape_clip():
    """BASELINE config C4: one 5-frame 480x832 clip, 2 instances - eval window of 3 and a training step."""
    m = _video_model(False)
    ev = _to_dev(synth.make_batch(b=1, n_f=3, n_i=2, H=480, W=832, edge_px=6.0, seed=9))
    with torch.no_grad():
        out = m(ev, mem_feat=None)
        out2 = m(ev, mem_feat=None, prev_pred=out["refined_masks"][:, 1])
    assert out["refined_masks"].shape == (1, 3, 2, 480, 832) and out["mem_feat"].shape == (1, 3, 128, 60, 104)
    assert bool(torch.isfinite(out2["refined_masks"]).all())
    m.train()
    tr = _to_dev(synth.make_batch(b=1, n_f=5, n_i=2, H=480, W=832, edge_px=6.0, seed=9, train=True, it=1))
    G.seed_all()
    o, loss = m(tr, mem_feat=None)
    (loss["total"] * 64.0).backward()
    assert o["temp_alpha"].shape[:2] == (1, 5) and np.isfinite(float(loss["total"]))
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)


def test_varying_batches_specialised_kernels_match_generic_kernels(monkeypatch):
    """Different batches (site counts, slot draws, ellipse sizes) through the routed kernels (K2b halo conv, K9b / K9c
    persistent sparse kernels) give the same losses and gradient norms as the generic kernels (K2, K9) they replace."""
    from maggie_b200 import _lib
    results = {}
    for mode in ("routed", "generic"):
        if mode == "generic":
            monkeypatch.setenv("MAGGIE_B200_NO_HALO_CONV", "1")
            monkeypatch.setenv("MAGGIE_B200_NO_PERSISTENT_SPARSE", "1")
        m = _model(True)
        m.decoder.inst_spec_layer.dropout.p = 0.0
        out = []
        for k, (seed, edge) in enumerate(((11, 3.0), (12, 8.0), (13, 5.0))):
            batch = _to_dev(synth.make_batch(b=4, n_f=1, n_i=3, H=256, W=256, edge_px=edge, seed=seed, train=True, it=1))
            for p in m.parameters():
                p.grad = None
            np.random.seed(100 + k)
            import random
            random.seed(100 + k)
            h0 = _lib.lib().mg_conv_halo_launches()
            _, loss = m(batch, mem_feat=None)
            (loss["total"] * 64.0).backward()
            gn = torch.stack([p.grad.float().norm() for p in m.parameters() if p.grad is not None])
            out.append((float(loss["total"]), gn, list(m.last_site_counts), _lib.lib().mg_conv_halo_launches() - h0))
        results[mode] = out
    for (l0, g0, c0, h0), (l1, g1, c1, h1) in zip(results["routed"], results["generic"]):
        assert c0 == c1 and c0[0] > 20000
        assert h0 > 0 and h1 == 0
        assert abs(l0 - l1) < 2e-2 * abs(l1), (l0, l1)
        rel = (g0 - g1).abs() / (g1.abs() + 1e-3 * g1.abs().max())
        assert float(rel.median()) < 2e-2, float(rel.median())
    assert len({tuple(c) for _, _, c, _ in results["routed"]}) == 3   # the three batches really differ
