"""-m gpu: the fp32-accurate evaluation mode (`MaGGIe.set_precision("high")`): split-fp16 operands on the tensor cores.

The north star asks for alpha within 1e-3 max-abs of the reference path.  The reference EVALUATES in fp32
(engine/test.py:131 has no autocast); fp16 storage of ~60 stacked layers costs 3e-3 .. 1e-2 on the random-weight test models
whatever the kernel (tools/precision_study.py reproduces that on the CPU), so the tolerance is met by the 'high' mode and
the fp16 mode's floor is asserted separately (tests/test_gpu_model.py).  Tolerances here are stated per test.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from maggie_b200 import _lib, dense, ops
from maggie_b200.config import CfgNode
from maggie_b200.network import build_model
from oracle import make_golden as G
from oracle import synth

pytestmark = pytest.mark.gpu
ALPHA_TOL = 1e-3          # north star: alpha within 1e-3 max abs
HIGH_FLOOR = 2.3e-4       # 1.25 x the measured worst case of the 'high' mode on the three eval goldens (1.84e-4, B200)


@pytest.fixture(autouse=True)
def _no_tf32():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def test_split_reconstructs_fp32():
    x = torch.randn(1 << 16, device="cuda") * torch.logspace(-6, 3, 1 << 16, device="cuda")
    hi, lo = dense.split_f32(x)
    assert hi.dtype == lo.dtype == torch.float16
    assert torch.equal(hi, x.half())
    err = (hi.double() + lo.double() - x.double()).abs()
    # 2^-22 relative, or the fp16 subnormal spacing (2^-24) / 2 absolute
    assert bool((err <= x.double().abs() * 2.0 ** -21 + 2.0 ** -25).all())


@pytest.mark.parametrize("shape", [
    # N, H, W, Ci, Co, k, stride, pad, dil
    (2, 32, 32, 32, 64, 3, 1, 1, 1), (2, 16, 16, 128, 128, 3, 1, 1, 1), (1, 16, 16, 64, 128, 3, 2, 1, 1),
    (2, 8, 8, 256, 256, 3, 1, 2, 2), (2, 16, 16, 64, 128, 2, 2, 0, 1), (2, 1, 1, 512, 256, 1, 1, 0, 1),
])
def test_conv_x3_matches_fp32(shape):
    N, H, W, Ci, Co, k, st, pad, dil = shape
    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    x = torch.randn(N, Ci, H, W, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    w = torch.randn(Co, Ci, k, k, device="cuda", generator=g) / (Ci * k * k) ** 0.5
    ref = F.conv2d(x.double(), w.double(), stride=st, padding=pad, dilation=dil)
    with torch.no_grad():
        y = dense.conv_bn_act(x, w, None, False, stride=st, padding=pad, dilation=dil, act=None)
    assert y.dtype == torch.float32
    err = float((y.double() - ref).abs().max())
    # fp32-level: the dropped lo*lo term is 2^-22 relative per product, the accumulator is fp32
    assert err < 2e-5 * float(ref.abs().max()) + 1e-6, err
    # the fp16 kernel on the same data is ~100x worse (this is what the mode buys)
    y16 = dense.conv_bn_act(x.half(), w, None, False, stride=st, padding=pad, dilation=dil, act=None)
    assert float((y16.double() - ref).abs().max()) > 10 * err


def test_conv_x3_epilogue_bn_residual_act_and_transposed():
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(2, 64, 16, 16, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    res = torch.randn(2, 64, 16, 16, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    w = torch.randn(64, 64, 3, 3, device="cuda", generator=g) / 24.0
    bn = torch.nn.BatchNorm2d(64).cuda().eval()
    with torch.no_grad():
        bn.running_mean.normal_(), bn.running_var.uniform_(0.5, 2.0), bn.weight.normal_(), bn.bias.normal_()
        y = dense.conv_bn_act(x, w, bn, False, act="lrelu", residual=res)
        ref = F.leaky_relu(bn(F.conv2d(x, w, padding=1)) + res, 0.2)
        assert float((y - ref).abs().max()) < 2e-5 * float(ref.abs().max())
        wt = torch.randn(64, 32, 4, 4, device="cuda", generator=g) / 32.0
        yt = dense.conv_bn_act(x, wt, None, False, act="relu", transposed=True)
        reft = F.relu(F.conv_transpose2d(x, wt, stride=2, padding=1))
        assert float((yt - reft).abs().max()) < 2e-5 * float(reft.abs().max())
        # low-resolution residual replicated 2x2 (decoder skip path)
        lo = torch.randn(2, 64, 8, 8, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
        yu = dense.conv_bn_act(x, w, None, False, act=None, residual=lo, res_up=True)
        refu = F.conv2d(x, w, padding=1) + F.interpolate(lo, scale_factor=2, mode="nearest")
        assert float((yu - refu).abs().max()) < 2e-5 * float(refu.abs().max())


def test_rows_ops_fp32():
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(2, 4096, 128, device="cuda", generator=g)
    w = torch.randn(128, 128, device="cuda", generator=g) / 11.3
    b = torch.randn(128, device="cuda", generator=g)
    with torch.no_grad():
        y = ops.linear_rows(x, w, b)
        ref = F.linear(x.double(), w.double(), b.double())
        assert y.dtype == torch.float32 and float((y.double() - ref).abs().max()) < 2e-5 * float(ref.abs().max())
        ln = torch.nn.LayerNorm(128).cuda()
        ln.weight.data.normal_(), ln.bias.data.normal_()
        r = torch.randn(2, 4096, 128, device="cuda", generator=g)
        z = ops.layer_norm(x, ln, residual=r)
        refz = F.layer_norm((x + r).double(), (128,), ln.weight.double(), ln.bias.double(), ln.eps)
        assert float((z.double() - refz).abs().max()) < 1e-5
        # attention cores: tokens <- pixels (with guidance statistic) and pixels <- tokens (with key padding)
        tok = torch.randn(2, 10, 128, device="cuda", generator=g)
        guid = torch.rand(2, 10, 4096, device="cuda", generator=g) > 0.7
        o, st = ops.attention(tok, x, r, None, guid)
        a = torch.softmax(torch.bmm(tok.double(), x.double().transpose(1, 2)) / 128 ** 0.5, -1)
        assert float((o.double() - torch.bmm(a, r.double())).abs().max()) < 2e-5
        assert float((st.double() - (a * guid).sum(-1)).abs().max()) < 2e-5
        pad = torch.zeros(2, 10, dtype=torch.bool, device="cuda")
        pad[:, 7:] = True
        o2, _ = ops.attention(x, tok, tok * 0.5, pad, None)
        s2 = (torch.bmm(x.double(), tok.double().transpose(1, 2)) / 128 ** 0.5).masked_fill(pad[:, None, :], float("-inf"))
        assert float((o2.double() - torch.bmm(torch.softmax(s2, -1), tok.double() * 0.5)).abs().max()) < 2e-5
        # OS8 head einsum
        f = torch.randn(2, 64, 32, 32, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
        t64 = torch.randn(2, 10, 64, device="cuda", generator=g)
        lg = ops.token_logits(t64, f, 1)
        refl = torch.einsum("bqc,bchw->bqhw", t64.double(), f.double())
        assert float((lg.double() - refl).abs().max()) < 2e-5 * float(refl.abs().max())


def _model(precision):
    m, _ = build_model(CfgNode(synth.model_cfg()))
    m.load_state_dict(synth.synth_state_dict(m.state_dict()), strict=True)
    return m.cuda().eval().set_precision(precision)


def _stage_errors(case, z, precision):
    """Runs the model on `case` and returns max-abs errors of the dense-stage tensors and outputs against the golden."""
    kw, _ = G.CASES[case]
    m = _model(precision)
    stage = m._stage[0]
    grabbed = {}
    aspp_fwd = stage.aspp.forward
    stage.aspp.forward = lambda x: grabbed.setdefault("aspp", aspp_fwd(x))
    imd_fwd = stage.decoder.refine_OS8.forward

    def imd(*a, **k):
        out = imd_fwd(*a, **k)
        grabbed["os8_logits"], grabbed["os8_feat"], grabbed["queries"] = out[0], out[1], out[2]
        return out

    stage.decoder.refine_OS8.forward = imd
    G.seed_all()
    batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in synth.make_batch(**kw).items()}
    with torch.no_grad():
        out = m(batch, mem_feat=None)
    err = {}
    for k in ("aspp", "os8_feat", "os8_logits", "queries"):
        ref = z["stage/" + k]
        got = grabbed[k].float().cpu().numpy()
        err[k] = (float(np.abs(got - ref).max()), float(np.abs(ref).max()))
    for k in ("alpha_os8", "alpha_os4", "alpha_os1", "refined_masks"):
        err[k] = (float(np.abs(out[k].float().cpu().numpy() - z["out/" + k]).max()), 1.0)
    err["detail_mask_agree"] = (float((out["detail_mask"].cpu().numpy() == z["out/detail_mask"]).mean()), 1.0)
    # The refinement region is a THRESHOLD of alpha_os8 (1/255 < a < 254/255, then dilated): a pixel whose alpha_os8 sits within
    # the fp32 noise of a threshold flips the detail mask there, the refined alpha at a flipped site is a different function
    # (OS8 value vs OS1 logit), and the 3x3 rulebook convs carry the difference a few pixels further.  Errors of the refined
    # outputs are therefore reported away from such flips (6 pixels; measured: 4 suffice) together with the number of flips.
    dis = torch.from_numpy((out["detail_mask"].cpu().numpy() != z["out/detail_mask"]).astype(np.float32))
    sh = dis.shape
    near = torch.nn.functional.max_pool2d(dis.reshape(-1, 1, *sh[-2:]), 13, 1, 6).reshape(sh).numpy() > 0
    err["flipped_px"] = (float(dis.sum()), float(dis.numel()))
    for k in ("alpha_os4", "alpha_os1", "refined_masks"):
        e = np.abs(out[k].float().cpu().numpy() - z["out/" + k])
        err[k + "_away_from_flips"] = (float(e[~near].max()), 1.0)
    return err


@pytest.mark.parametrize("case", [c for c in G.CASES if c.startswith("eval")])
def test_high_precision_eval_meets_the_north_star_tolerance(case, golden):
    z = golden(case)
    hi = _stage_errors(case, z, "high")
    lo = _stage_errors(case, z, "fp16")
    print(f"\n{case}: max-abs error vs the fp32 reference (|ref|max)  high / fp16")
    for k in hi:
        print(f"  {k:18s} {hi[k][0]:.3e} / {lo[k][0]:.3e}   ({hi[k][1]:.3g})")
    assert hi["alpha_os8"][0] <= ALPHA_TOL, hi["alpha_os8"]
    assert hi["alpha_os8"][0] <= HIGH_FLOOR, hi["alpha_os8"]          # 1.25x-of-measured style bound (see module docstring)
    for k in ("aspp", "os8_feat", "os8_logits", "queries"):
        assert hi[k][0] <= 2e-4 * max(1.0, hi[k][1]), (k, hi[k])
    assert hi["detail_mask_agree"][0] >= 0.9999
    # every alpha the model returns - OS4, OS1 and the fused result - meets the tolerance away from the (<= 8 per case:
    # measured 2 / 8 / 1) threshold flips of the detail mask; measured 3.7e-4 .. 4.5e-4
    assert hi["flipped_px"][0] <= 16 and hi["flipped_px"][0] <= 1e-4 * hi["flipped_px"][1]
    for k in ("alpha_os4", "alpha_os1", "refined_masks"):
        assert hi[k + "_away_from_flips"][0] <= ALPHA_TOL, (k, hi[k + "_away_from_flips"])
        assert hi[k + "_away_from_flips"][0] <= 1.25 * 4.5e-4, (k, hi[k + "_away_from_flips"])
    # and the mode is what makes the difference: fp16 storage is at least 10x further away on the same case
    assert lo["alpha_os8"][0] > 10 * hi["alpha_os8"][0]
    assert _lib.launch_count() > 0


def test_high_precision_is_eval_only_and_image_only():
    m = _model("high")
    assert m.precision == "high"
    with pytest.raises(ValueError):
        m.set_precision("fp64")
    vm, _ = build_model(CfgNode(synth.video_cfg()))
    with pytest.raises(NotImplementedError):
        vm.set_precision("high")
    # training ignores the mode (fp16 path)
    m.train()
    tb = synth.make_batch(b=2, n_f=1, n_i=2, H=128, W=128, edge_px=4.0, seed=6, train=True, it=1)
    G.seed_all()
    _, loss = m({k: (v.cuda() if torch.is_tensor(v) else v) for k, v in tb.items()}, mem_feat=None)
    assert np.isfinite(float(loss["total"]))
