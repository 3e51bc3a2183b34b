"""CPU tests of the loader / evaluator ends (K16): known answers for the oracle restatement and the no-fallback rule."""
import numpy as np
import pytest
import torch

from oracle import io_oracle as O

MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def test_input_oracle_known_answers():
    fr = torch.zeros((1, 16, 16, 3), dtype=torch.uint8)
    fr[0, 0, 0] = torch.tensor([255, 0, 128], dtype=torch.uint8)
    al = torch.zeros((1, 2, 16, 16), dtype=torch.uint8)
    al[0, 0, 0, :4] = torch.tensor([4, 5, 255, 128], dtype=torch.uint8)
    mk = torch.zeros((1, 2, 16, 16), dtype=torch.uint8)
    mk[0, 1, 8, 8] = 255
    mk[0, 1, 9, 9] = 255          # not on the 1/8 grid: dropped by the nearest down-sampling
    img, alpha, mask = O.to_model_inputs(fr, al, mk, MEAN, STD)
    assert img.shape == (1, 3, 16, 16) and alpha.shape == (1, 2, 16, 16) and mask.shape == (1, 2, 2, 2)
    want = [(1.0 - MEAN[0]) / STD[0], (0.0 - MEAN[1]) / STD[1], (128 / 255 - MEAN[2]) / STD[2]]
    assert np.allclose(img[0, :, 0, 0].numpy(), want, atol=1e-6)
    assert alpha[0, 0, 0, :4].tolist() == pytest.approx([0.0, 5 / 255, 1.0, 128 / 255], abs=1e-7)
    assert mask[0, 1].tolist() == [[0.0, 0.0], [0.0, 1.0]] and float(mask[0, 0].sum()) == 0.0
    full = O.to_model_inputs(fr, al, mk, MEAN, STD, downscale_mask=False)[2]
    assert full.shape == (1, 2, 16, 16) and float(full.sum()) == 2.0


def test_finalize_oracle_known_answers():
    a = torch.linspace(0, 1, 6 * 8).reshape(1, 1, 6, 8)
    out, _ = O.finalize_alpha(a, [{"name": "resize", "ori_size": (4, 6)}, {"name": ["padding"], "pad_size": (2, 2)}])
    assert out.shape == (1, 1, 4, 6)          # crop to 4 x 6, then an identity resize
    crop = a[..., :4, :6].numpy().copy()
    crop[crop <= 1 / 255] = 0
    crop[crop >= 254 / 255] = 1
    assert np.allclose(out, crop, atol=1e-6)
    up, _ = O.finalize_alpha(a, [{"name": "resize", "ori_size": (11, 15)}])
    assert up.shape == (1, 1, 11, 15)
    assert up[0, 0, 0, 0] == 0.0 and up[0, 0, -1, -1] == 1.0      # align_corners: corners map to corners (then clamped)
    assert np.all(np.diff(up[0, 0, 5]) >= 0)


def test_product_path_has_no_cpu_fallback():
    from maggie_b200 import io
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        io.prepare_batch(torch.zeros((1, 1, 8, 8, 3), dtype=torch.uint8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        io.finalize_alpha(torch.zeros((1, 1, 1, 8, 8)))


def _soft_blobs(n, H, W, seed):
    """uint8 alpha planes with soft-edged blobs (values over the whole 0..255 range) + hard masks that disagree in places."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    al = np.zeros((n, H, W), np.float32)
    for i in range(n):
        for _ in range(3):
            cy, cx, r = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(3, min(H, W) / 3)
            al[i] = np.maximum(al[i], np.clip((r - np.hypot(yy - cy, xx - cx)) / 4.0 + 0.5, 0, 1))
    al = (al * 255).round().astype(np.uint8)
    mk = ((al > 100) * 255).astype(np.uint8)
    return al, mk


@pytest.mark.parametrize("k,iterations", [(25, 1), (2, 5), (3, 7), (4, 14), (1, 1), (7, 2)])
def test_transition_oracle_is_pinned_to_opencv(k, iterations):
    """oracle/io_oracle.py::transition_gt against the reference's own expression sequence on cv2 (dataloader/utils.py:15-35):
    cv2.getStructuringElement(MORPH_ELLIPSE) + cv2.dilate / cv2.erode with `iterations`, the `> 0` test and the mask
    disagreement, masks at full size and at 1/8 (repeat_interleave)."""
    cv2 = pytest.importorskip("cv2")
    al, mk = _soft_blobs(3, 48, 64, seed=k * 31 + iterations)
    mk8 = mk[:, ::8, ::8].copy()
    kernel = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))
    for masks in (None, mk, mk8):
        ref = []
        for x in torch.from_numpy(al)[:, None]:
            dil = cv2.dilate(x[0, :, :, None].numpy(), kernel, iterations=iterations)
            ero = cv2.erode(x[0, :, :, None].numpy(), kernel, iterations=iterations)
            ref.append(torch.from_numpy(((dil.astype(np.float32) - ero.astype(np.float32)) > 0).astype(float)))
        ref = torch.stack(ref).unsqueeze(1)
        if masks is not None:
            m = torch.from_numpy(masks)[:, None]
            if m.shape[-1] != al.shape[-1]:
                m = torch.repeat_interleave(torch.repeat_interleave(m, 8, dim=-1), 8, dim=-2)
            ref[(torch.from_numpy(al)[:, None] > 127) != (m == 255)] = 1.0
        got = O.transition_gt(al, masks, k_size=k, iterations=iterations)
        assert np.array_equal(got, ref[:, 0].numpy().astype(np.uint8)), (k, iterations, masks is not None)
