"""K16 against the oracle restatement of the reference's loader tail (ToTensor + Normalize + dataset scaling) and
evaluation tail (reverse_transform_tensor + clamps): integer-derived values bit-exact, the bilinear resize to 2e-6."""
import numpy as np
import pytest
import torch

from oracle import io_oracle as O

pytestmark = pytest.mark.gpu
MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


@pytest.mark.parametrize("b,n_f,n_i,H,W,down", [(2, 1, 3, 64, 96, True), (1, 3, 2, 72, 40, False), (1, 1, 1, 8, 8, True)])
def test_input_stage_bit_exact(b, n_f, n_i, H, W, down):
    from maggie_b200 import io
    g = torch.Generator().manual_seed(3)
    fr = torch.randint(0, 256, (b, n_f, H, W, 3), generator=g, dtype=torch.uint8)
    al = torch.randint(0, 256, (b, n_f, n_i, H, W), generator=g, dtype=torch.uint8)
    al[..., :4] = torch.randint(0, 8, al[..., :4].shape, generator=g, dtype=torch.uint8)     # exercise the `< 5` rule
    mk = (torch.randint(0, 2, (b, n_f, n_i, H, W), generator=g, dtype=torch.uint8) * 255)
    out = io.prepare_batch(fr.cuda(), al.cuda(), mk.cuda(), MEAN, STD, downscale_mask=down)
    img, alpha, mask = O.to_model_inputs(fr.reshape(b * n_f, H, W, 3), al.reshape(b * n_f, n_i, H, W),
                                         mk.reshape(b * n_f, n_i, H, W), MEAN, STD, downscale_mask=down)
    assert out["image"].shape == (b, n_f, 3, H, W)
    assert torch.equal(out["image"].cpu().reshape(img.shape), img)
    assert torch.equal(out["alpha"].cpu().reshape(alpha.shape), alpha)
    assert torch.equal(out["mask"].cpu().reshape(mask.shape), mask)
    only = io.prepare_batch(fr.cuda())
    assert set(only) == {"image"} and torch.equal(only["image"], out["image"])


@pytest.mark.parametrize("info", [
    [],
    [{"name": ["padding"], "pad_size": (torch.tensor(5), torch.tensor(12))}],
    [{"name": "resize", "ori_size": (181, 263)}],
    [{"name": ["resize"], "ori_size": (torch.tensor(300), torch.tensor(150))}, {"name": ["padding"], "pad_size": (7, 0)}],
    [{"name": "resize", "ori_size": (64, 64)}, {"name": "padding", "pad_size": (0, 3)}, {"name": "resize", "ori_size": (90, 70)}],
])
def test_alpha_finalize(info):
    from maggie_b200 import io
    g = torch.Generator().manual_seed(11)
    a = torch.rand((1, 2, 3, 96, 128), generator=g)
    a[..., :20, :] = 0.0
    a[..., 60:, :] = 1.0
    a[..., 30:40, :] = 0.5 / 255.0
    plain = [{k: ([x.item() if torch.is_tensor(x) else x for x in v] if k != "name" else v) for k, v in t.items()} for t in info]
    want, pre = O.finalize_alpha(a, plain)
    got = io.finalize_alpha(a.cuda(), info).cpu().numpy()
    assert got.shape == want.shape
    # away from the two clamp thresholds the results agree to rounding; at a threshold a last-bit difference may flip
    safe = (np.abs(pre - 1.0 / 255.0) > 1e-5) & (np.abs(pre - 254.0 / 255.0) > 1e-5)
    assert safe.mean() > 0.99
    assert np.abs(got - want)[safe].max() <= 2e-6
    assert ((got == 0.0) | (got > 1.0 / 255.0 - 1e-6)).all() and ((got == 1.0) | (got < 254.0 / 255.0 + 1e-6)).all()
