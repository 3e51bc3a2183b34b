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
