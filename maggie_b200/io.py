"""The two ends of the hot path that touch loader / evaluator data, on the GPU (K16).

`prepare_batch` replaces the float conversion at the end of the reference's data pipeline - `ToTensor` + `Normalize`
(dataloader/transforms.py:720-783) and the scaling / mask down-sampling of the datasets (dataloader/him.py:156-157,
175-176) - so that the loader ships uint8 (1 byte per value over PCIe instead of 4) and never builds full-size float
tensors on the CPU.  `finalize_alpha` replaces `reverse_transform_tensor` (utils/postprocessing.py:36-64) plus the
near-0 / near-1 clamps of the evaluation loop (engine/test.py:141-142).  Both need the CUDA library: no CPU fallback."""
import ctypes

import torch

from . import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)     # configs/maggie_image.yaml `dataset.*.mean / std` defaults of the reference
IMAGENET_STD = (0.229, 0.224, 0.225)


def _f3(v):
    return (ctypes.c_float * 3)(*[float(x) for x in v])


def prepare_batch(frames, alphas=None, masks=None, mean=IMAGENET_MEAN, std=IMAGENET_STD, downscale_mask=True):
    """frames uint8 [b, n_f, H, W, 3] (as decoded); alphas / masks uint8 [b, n_f, n_i, H, W] (0..255) or None.
    -> dict(image [b,n_f,3,H,W] fp32 normalised, alpha [b,n_f,n_i,H,W] fp32 in [0,1] with values < 5/255 zeroed,
    mask [b,n_f,n_i,H/8,W/8] (downscale_mask) or full size, fp32) - the tensors `MaGGIe.forward` takes."""
    _lib.need_cuda(frames, alphas, masks)
    if frames.dtype != torch.uint8 or frames.dim() != 5 or frames.shape[-1] != 3:
        raise ValueError("prepare_batch: frames must be uint8 [b, n_f, H, W, 3]")
    b, n_f, H, W, _ = frames.shape
    n_i = 0
    for t in (alphas, masks):
        if t is not None:
            if t.dtype != torch.uint8 or t.dim() != 5 or t.shape[:2] != (b, n_f) or t.shape[-2:] != (H, W):
                raise ValueError("prepare_batch: alphas / masks must be uint8 [b, n_f, n_i, H, W]")
            n_i = t.shape[2]
    if alphas is not None and masks is not None and alphas.shape != masks.shape:
        raise ValueError("prepare_batch: alphas and masks differ in shape")
    div = 8 if downscale_mask else 1
    if masks is not None and div == 8 and (H % 8 or W % 8):
        raise ValueError("prepare_batch: mask down-sampling needs H, W multiples of 8")
    dev = frames.device
    out = {"image": torch.empty((b, n_f, 3, H, W), dtype=torch.float32, device=dev)}
    if alphas is not None:
        out["alpha"] = torch.empty((b, n_f, n_i, H, W), dtype=torch.float32, device=dev)
    if masks is not None:
        out["mask"] = torch.empty((b, n_f, n_i, H // div, W // div), dtype=torch.float32, device=dev)
    p = _lib.tensor_ptr
    _lib.check(_lib.lib().mg_input_stage(
        p(frames.contiguous()), p(alphas.contiguous()) if alphas is not None else None,
        p(masks.contiguous()) if masks is not None else None, p(out["image"]), p(out.get("alpha")), p(out.get("mask")),
        _f3(mean), _f3(std), b * n_f, n_i, H, W, div, _lib.stream_ptr()), "mg_input_stage")
    return out


def finalize_alpha(alpha, transform_info=(), lo=1.0 / 255.0, hi=254.0 / 255.0):
    """alpha [..., h, w] fp32 (e.g. `output['refined_masks']`) -> [..., H, W]: the test transforms undone in reverse order
    (`{'name': 'padding', 'pad_size': (ph, pw)}` crops, `{'name': 'resize', 'ori_size': (H, W)}` resizes bilinearly with
    align_corners = True) and values <= lo / >= hi snapped to 0 / 1."""
    _lib.need_cuda(alpha)
    if alpha.dtype != torch.float32:
        raise ValueError("finalize_alpha: alpha must be fp32")
    lead = tuple(alpha.shape[:-2])
    cur = alpha.contiguous().reshape(-1, *alpha.shape[-2:])
    steps = []
    for tr in list(transform_info)[::-1]:
        name = tr["name"][0] if isinstance(tr["name"], (list, tuple)) else tr["name"]
        if name == "padding":
            steps.append(("pad", *[int(v) for v in tr["pad_size"]]))
        elif name == "resize":
            steps.append(("resize", *[int(v) for v in tr["ori_size"]]))
    # (crop, resize) pairs fuse into one launch; the clamp runs in the last one
    launches, i = [], 0
    while i < len(steps):
        if steps[i][0] == "pad" and i + 1 < len(steps) and steps[i + 1][0] == "resize":
            launches.append((steps[i][1], steps[i][2], steps[i + 1][1], steps[i + 1][2]))
            i += 2
        elif steps[i][0] == "pad":
            launches.append((steps[i][1], steps[i][2], 0, 0))
            i += 1
        else:
            launches.append((0, 0, steps[i][1], steps[i][2]))
            i += 1
    if not launches:
        launches = [(0, 0, 0, 0)]
    L = _lib.lib()
    for k, (ph, pw, oh, ow) in enumerate(launches):
        h, w = cur.shape[-2:]
        Ho, Wo = (oh, ow) if oh > 0 else (h - ph, w - pw)
        nxt = torch.empty((cur.shape[0], Ho, Wo), dtype=torch.float32, device=cur.device)
        last = k == len(launches) - 1
        _lib.check(L.mg_alpha_finalize(_lib.tensor_ptr(cur), _lib.tensor_ptr(nxt), cur.shape[0], h, w, ph, pw, oh, ow,
                                       float(lo) if last else -1e30, float(hi) if last else 1e30, _lib.stream_ptr()),
                   "mg_alpha_finalize")
        cur = nxt
    return cur.reshape(*lead, *cur.shape[-2:])


def transition_gt(alphas, masks=None, k_size=25, iterations=1, temporal=False):
    """`gen_transition_gt` / `gen_transition_temporal_gt` of the loaders (dataloader/utils.py:15-60) on the GPU (K18).
    alphas uint8 [n, H, W] (one plane per instance, or per frame for the temporal form); masks uint8 [n, H, W] or
    [n, H/8, W/8] or None -> uint8 {0,1} [n, H, W].  temporal: planes i >= 1 are additionally zeroed where
    alphas[i] - alphas[i-1] <= 1/255 (the reference compares the raw uint8 tensors: a non-positive difference), before the
    mask disagreement is OR-ed in."""
    _lib.need_cuda(alphas, masks)
    if alphas.dtype != torch.uint8 or alphas.dim() != 3:
        raise ValueError("transition_gt: alphas must be uint8 [n, H, W]")
    n, H, W = alphas.shape
    div = 1
    if masks is not None:
        if masks.dtype != torch.uint8 or masks.dim() != 3 or masks.shape[0] != n:
            raise ValueError("transition_gt: masks must be uint8 [n, h, w]")
        if tuple(masks.shape[1:]) == (H, W):
            div = 1
        elif tuple(masks.shape[1:]) == (H // 8, W // 8) and H % 8 == 0 and W % 8 == 0:
            div = 8
        else:
            raise ValueError("transition_gt: masks must be full size or 1/8 size")
    a = alphas.contiguous()
    out = torch.empty_like(a)
    tmp = torch.empty(4 * a.numel(), dtype=torch.uint8, device=a.device) if iterations > 1 else None
    p = _lib.tensor_ptr
    run = lambda m, o: _lib.check(_lib.lib().mg_transition_gt(p(a), p(m), div, n, H, W, int(k_size), int(iterations), p(tmp), p(o),
                                                              _lib.stream_ptr()), "mg_transition_gt")
    if not temporal:
        run(masks.contiguous() if masks is not None else None, out)
        return out
    run(None, out)
    if n > 1:
        sparse = (a[1:].float() - a[:-1].float()) > 1.0 / 255.0
        out[1:] = out[1:] * sparse.to(torch.uint8)
    if masks is not None:
        up = masks if div == 1 else masks.repeat_interleave(8, -1).repeat_interleave(8, -2)
        out = torch.where((a > 127) != (up == 255), torch.ones_like(out), out)
    return out
