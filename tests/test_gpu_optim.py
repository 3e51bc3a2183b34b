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
