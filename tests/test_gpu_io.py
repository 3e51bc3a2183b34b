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


# ---- K18: transition ("trimap") ground truth ------------------------------------------------------------------------
def _soft_blobs(n, H, W, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    al = np.zeros((n, H, W), np.float32)
    for i in range(n):
        for _ in range(3):
            cy, cx, r = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(3, min(H, W) / 3)
            al[i] = np.maximum(al[i], np.clip((r - np.hypot(yy - cy, xx - cx)) / 4.0 + 0.5, 0, 1))
    al = (al * 255).round().astype(np.uint8)
    return al, ((al > 100) * 255).astype(np.uint8)


@pytest.mark.parametrize("k,iterations,H,W", [(25, 1, 96, 160), (2, 5, 64, 72), (3, 9, 40, 104), (4, 14, 72, 64), (1, 1, 32, 32),
                                              (29, 1, 50, 70), (7, 3, 33, 45)])
def test_transition_gt_bit_exact(k, iterations, H, W):
    """K18 against the oracle restatement of gen_transition_gt (itself pinned to OpenCV in tests/test_io_host.py): bit exact,
    without masks, with full-size masks and with the loader's 1/8 masks."""
    from maggie_b200 import io as mio
    from oracle import io_oracle as O
    al, mk = _soft_blobs(3, H, W, seed=k + 7 * iterations)
    cases = [None, mk] + ([mk[:, ::8, ::8].copy()] if H % 8 == 0 and W % 8 == 0 else [])
    for masks in cases:
        ref = O.transition_gt(al, masks, k_size=k, iterations=iterations)
        got = mio.transition_gt(torch.from_numpy(al).cuda(), None if masks is None else torch.from_numpy(masks).cuda(),
                                k_size=k, iterations=iterations)
        assert got.dtype == torch.uint8 and np.array_equal(got.cpu().numpy(), ref), (k, iterations, masks is not None)


def test_transition_temporal_gt_matches_the_reference_sequence():
    """gen_transition_temporal_gt (dataloader/utils.py:37-60): planes i >= 1 zeroed where the alpha did not grow by more than
    1/255 from the previous frame, then the mask disagreement."""
    from maggie_b200 import io as mio
    from oracle import io_oracle as O
    al, mk = _soft_blobs(4, 64, 96, seed=3)
    mk8 = mk[:, ::8, ::8].copy()
    ref = O.transition_gt(al, None, k_size=25, iterations=1).astype(np.float32)
    a = torch.from_numpy(al)[:, None]
    sparsity = ((a[1:].float() - a[:-1].float()) > 1.0 / 255.0).float()       # the reference subtracts the float-converted tensors
    for i in range(1, 4):
        ref[i][sparsity[i - 1, 0].numpy() == 0] = 0.0
    up = np.repeat(np.repeat(mk8, 8, axis=-1), 8, axis=-2)
    ref[(al > 127) != (up == 255)] = 1.0
    got = mio.transition_gt(torch.from_numpy(al).cuda(), torch.from_numpy(mk8).cuda(), temporal=True)
    assert np.array_equal(got.cpu().numpy(), ref.astype(np.uint8))
