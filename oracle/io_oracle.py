"""TEST INFRASTRUCTURE ONLY - CPU restatement of the loader / evaluator ends of the path (checker for K16).

Pinning: every function below is the reference's own torch expression sequence, applied on the CPU (the reference
modules themselves cannot be imported here: dataloader/transforms.py needs imgaug, utils/postprocessing.py needs
scikit-image, both absent from this image) - the arithmetic is a handful of torch calls, restated one to one:
  * to_model_inputs : ToTensor + Normalize, dataloader/transforms.py:720-783 (`permute(0,3,1,2).float()`,
    `alphas[alphas < 5] = 0`, `frames / 255.0`, `(frames - mean) / std`), dataset scaling dataloader/him.py:156-157
    (`alpha * 1.0 / 255`, `mask * 1.0 / 255`) and mask down-sampling him.py:175-176 (`F.interpolate(..., mode="nearest")`).
  * finalize_alpha  : utils/postprocessing.py:36-64 (`reverse_transform_tensor`: crop `img[:, :h-pad_h, :w-pad_w]`,
    `F.interpolate(..., mode='bilinear', align_corners=True)`) and engine/test.py:141-142 (clamps)."""
import numpy as np
import torch
import torch.nn.functional as F


def to_model_inputs(frames, alphas, masks, mean, std, downscale_mask=True):
    """frames uint8 [T,H,W,3]; alphas / masks uint8 [T,n_i,H,W] -> image [T,3,H,W], alpha, mask (fp32)."""
    fr = frames.permute(0, 3, 1, 2).contiguous().float()
    fr = fr / 255.0
    fr = (fr - torch.tensor(mean).view(1, 3, 1, 1).float()) / torch.tensor(std).view(1, 3, 1, 1).float()
    a = alphas.clone()
    a[a < 5] = 0
    alpha = (a * 1.0 / 255).float()
    mask = (masks * 1.0 / 255).float()
    if downscale_mask:
        mask = F.interpolate(mask, size=(fr.shape[2] // 8, fr.shape[3] // 8), mode="nearest")
    return fr, alpha, mask


def finalize_alpha(img, transform_info):
    """img [..., h, w] fp32 -> numpy [..., H, W] after the reversed transforms and the clamps."""
    shape = list(img.shape)
    img = img.reshape(-1, *img.shape[-2:])
    for tr in list(transform_info)[::-1]:
        name = tr["name"][0] if isinstance(tr["name"], list) else tr["name"]
        if name == "padding":
            ph, pw = tr["pad_size"]
            h, w = img.shape[-2:]
            img = img[:, :h - ph, :w - pw]
        elif name == "resize":
            h, w = tr["ori_size"]
            img = F.interpolate(img.unsqueeze(1), size=(h, w), mode="bilinear", align_corners=True).squeeze(1)
    shape[-2:] = img.shape[-2:]
    pre = img.reshape(shape).numpy().copy()
    out = pre.copy()
    out[out <= 1.0 / 255.0] = 0.0
    out[out >= 254.0 / 255.0] = 1.0
    return out, pre


def _morph(x, k, iterations, dilate):
    """Grey-level dilation / erosion of a uint8 [H, W] plane by cv2's MORPH_ELLIPSE (k, k), anchor (k // 2, k // 2), pixels outside
    the image ignored (cv2's default morphology border), `iterations` times.  numpy restatement, pinned against cv2 itself in
    tests/test_io_host.py."""
    from .unknown import ellipse_spans
    H, W = x.shape
    a = k // 2
    spans = ellipse_spans(k)
    neutral = 0 if dilate else 255
    cur = x.astype(np.uint8)
    for _ in range(iterations):
        pad = np.full((H + 2 * k, W + 2 * k), neutral, np.uint8)
        pad[k:k + H, k:k + W] = cur
        out = np.full((H, W), neutral, np.uint8)
        for i, (j1, j2) in enumerate(spans):
            for j in range(j1, j2):
                win = pad[k + i - a:k + i - a + H, k + j - a:k + j - a + W]
                out = np.maximum(out, win) if dilate else np.minimum(out, win)
        cur = out
    return cur


def transition_gt(alphas, masks=None, k_size=25, iterations=1):
    """dataloader/utils.py:15-35 on uint8 numpy planes [n, H, W] (masks [n, H, W] or [n, H/8, W/8]) -> uint8 {0,1} [n, H, W]."""
    out = []
    for x in alphas:
        d = _morph(x, k_size, iterations, True).astype(np.int32)
        e = _morph(x, k_size, iterations, False).astype(np.int32)
        out.append(((d - e) > 0).astype(np.uint8))
    out = np.stack(out)
    if masks is not None:
        if masks.shape[-1] != alphas.shape[-1]:
            masks = np.repeat(np.repeat(masks, 8, axis=-1), 8, axis=-2)
        out[(alphas > 127) != (masks == 255)] = 1
    return out
