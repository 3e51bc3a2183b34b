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
