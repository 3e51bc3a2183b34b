"""-m gpu: K14 fused optimizer tail against torch (manual unscale + clip_grad_norm_(0.01) + torch.optim.AdamW), including a
step whose gradient contains an inf (skipped, like GradScaler.step)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_adamw_matches_torch_unscale_clip_adamw():
    from maggie_b200.dp import FlatGradAllReduce
    from maggie_b200.optim import FusedAdamW
    torch.manual_seed(0)
    dev = torch.device("cuda")
    shapes = [(64, 32, 3, 3), (17,), (128, 128), (1,), (40000,), (3, 5, 7)]
    ours = [torch.nn.Parameter(torch.randn(s, device=dev)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    flat = FlatGradAllReduce(ours)
    kw = dict(lr=1.5e-4, betas=(0.5, 0.999), weight_decay=0.01)
    opt = FusedAdamW(flat, clip_norm=0.01, **kw)
    ropt = torch.optim.AdamW(ref, **kw)
    scale = 128.0
    for it in range(5):
        lr = 1.5e-4 * (1 + it)                      # a scheduler changing the rate between steps
        opt.param_groups[0]["lr"] = lr
        ropt.param_groups[0]["lr"] = lr
        grads = [torch.randn_like(p) * (0.001 if it % 2 else 3.0) for p in ours]     # below and above the clip threshold
        if it == 2:
            grads[2][5, 7] = float("inf")
        flat.zero()
        for p, g in zip(ours, grads):
            p.grad = g * scale                      # what a backward of (loss * scale) leaves behind
        rep = opt.step(grad_scale=scale).clone()
        if it == 2:
            assert float(rep[1]) == 1.0
        else:
            assert float(rep[1]) == 0.0
            for p, g in zip(ref, grads):
                p.grad = g.clone()
            norm = torch.nn.utils.clip_grad_norm_(ref, 0.01)
            ropt.step()
            assert abs(float(rep[0]) - float(norm)) < 1e-4 * float(norm)
        for a, b in zip(ours, ref):
            assert torch.allclose(a, b, rtol=2e-5, atol=2e-7), (it, float((a - b).abs().max()))
    assert float(opt.step_count) == 4.0            # the inf step did not count


def test_parameters_without_gradient_are_left_alone_and_state_dict_is_torch_format():
    """torch.optim.AdamW skips `grad is None` parameters entirely (no weight decay, no moment decay) - the four
    `dummy_downscale` weights of the model are such parameters; and the optimizer checkpoint (`last_opt.pth` of the reference
    engine) is interchangeable with torch's."""
    from maggie_b200.dp import FlatGradAllReduce
    from maggie_b200.optim import FusedAdamW
    torch.manual_seed(1)
    dev = torch.device("cuda")
    shapes = [(32, 16), (9,), (5, 5, 5), (1000,)]
    ours = [torch.nn.Parameter(torch.randn(s, device=dev)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    flat = FlatGradAllReduce(ours)
    kw = dict(lr=1e-2, betas=(0.9, 0.999), weight_decay=0.1)
    opt, ropt = FusedAdamW(flat, clip_norm=1e9, **kw), torch.optim.AdamW(ref, **kw)
    frozen = ours[2].detach().clone()
    for it in range(3):
        grads = [torch.randn_like(p) for p in ours]
        flat.zero()
        for i, (p, r, g) in enumerate(zip(ours, ref, grads)):
            r.grad = None
            if i != 2:                               # parameter 2 never receives a gradient
                p.grad, r.grad = g.clone(), g.clone()
        opt.step()
        ropt.step()
    assert torch.equal(ours[2], frozen), "a parameter without gradient was decayed"
    for a, b in zip(ours, ref):
        assert torch.allclose(a, b, rtol=2e-5, atol=2e-7)
    # torch AdamW can resume from our checkpoint and vice versa
    sd = opt.state_dict()
    assert set(sd["state"]) == {0, 1, 3} and set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    fresh = torch.optim.AdamW([torch.nn.Parameter(p.detach().clone()) for p in ours], **kw)
    fresh.load_state_dict(sd)
    rsd = ropt.state_dict()
    for i in (0, 1, 3):
        assert torch.allclose(sd["state"][i]["exp_avg"], rsd["state"][i]["exp_avg"], rtol=1e-5, atol=1e-8)
        assert torch.allclose(sd["state"][i]["exp_avg_sq"], rsd["state"][i]["exp_avg_sq"], rtol=1e-4, atol=1e-10)
        assert float(sd["state"][i]["step"]) == float(rsd["state"][i]["step"]) == 3.0
    opt2 = FusedAdamW(FlatGradAllReduce(ours), clip_norm=1e9, **kw)
    opt2.load_state_dict(rsd)
    assert torch.allclose(opt2.m, opt.m, rtol=1e-4, atol=1e-8) and float(opt2.step_count) == 3.0


def test_bank_gradients_land_in_the_flat_buffer_without_a_pack():
    """K0's grouped backward writes the conv-weight gradients straight into the head of the flat buffer
    (`FlatGradAllReduce(..., bank=model.bank)`): after a backward those p.grad ARE the flat views, the values equal a run
    without the hand-over, and a second backward without zero() still accumulates like torch."""
    import random
    import numpy as np
    from maggie_b200.config import CfgNode
    from maggie_b200.dp import FlatGradAllReduce
    from maggie_b200.network import build_model
    from oracle import synth
    dev = torch.device("cuda")
    torch.manual_seed(5)
    model, _ = build_model(CfgNode(synth.model_cfg()))
    model.to(dev).train()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    batch = synth.make_batch(b=2, n_f=1, n_i=2, H=128, W=128, edge_px=4.0, seed=6, train=True, it=1)
    batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}

    def backward(scale=128.0):
        np.random.seed(7), random.seed(7), torch.manual_seed(7)
        model.load_state_dict(sd)
        _, loss = model(batch, mem_feat=None)
        (loss["total"] * scale).backward()

    def rel(a, b):
        return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))

    for p in model.parameters():
        p.grad = None
    backward()
    want = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    flat = FlatGradAllReduce(model.parameters(), bank=model.bank)
    flat.zero()
    backward()
    bank_ws = {id(e.w) for e in model.bank.entries}
    n_direct = 0
    for p, v in zip(flat.params, flat.views):
        if id(p) in bank_ws:
            assert p.grad is not None and p.grad.data_ptr() == v.data_ptr()
            n_direct += p.numel()
    assert n_direct > 0.9 * flat.flat.numel()
    flat.pack()
    # (the training backward is not bit-reproducible - fp32 atomics in the BatchNorm statistics - hence norms, not bits)
    # and the ASPP pooling branch normalises over the 2 frames of this batch: a few gradients move by several per cent
    # between two runs from identical state; the median is stable)
    med = lambda k: float(np.median([rel(p.grad, k * want[n]) for n, p in model.named_parameters() if n in want and id(p) in bank_ws]))
    assert med(1) < 5e-2, med(1)
    backward(3 * 128.0)                            # no zero(): torch semantics = accumulate -> g + 3 g (an overwrite: 6 g)
    assert med(4) < 5e-2, med(4)
