"""Training losses evaluated inside forward(): weighted L1, 3-level Laplacian pyramid, Sobel gradient at the three
alpha scales.  Reference: arch/maggie.py:237-368 (regression_loss, compute_loss), loss.py:67-191 (GradientLoss,
LapLoss).  The heavy part - every per-pixel stencil and reduction, forward and backward - is ONE native op
(`ops.matte_loss_sums`, K12); this file only normalises the 24 partial sums into the reference's loss dictionary.
"""
from .. import ops


def compute_loss(pred, w4, w1, alphas, cfg, valid=None):
    """Image-model loss dictionary (dtSSD is handled by the video subclass).  `valid` [B, planes, 1, 1] {0,1}: the
    predictions are the UNMASKED alphas and the reference's `pred * valid_masks` (arch/maggie.py:112-117) is applied inside
    the loss kernels (three full-size products and their backward less)."""
    a1, a4, a8 = pred["alpha_os1"], pred["alpha_os4"], pred["alpha_os8"]
    w8 = (alphas.sum((2, 3), keepdim=True) > 0).to(a8.dtype).expand_as(a8)
    if cfg.loss_reweight_os8:
        lo, hi = 1.0 / 255.0, 254.0 / 255.0
        unk_pred = (a8 <= hi) & (a8 >= lo)
        if valid is not None:
            unk_pred = unk_pred & (valid > 0)       # a masked-out plane is identically 0: never "unknown"
        unk = ((alphas <= hi) & (alphas >= lo)) | unk_pred
        w8 = unk.to(a8.dtype) + w8
    # sums[scale] = [sum|p w - t w|, sum|L_0| w_0, sum|L_1| w_1, sum|L_2| w_2, sum|sobel(pw) - sobel(tw)|, sum w_0, sum w_1, sum w_2]
    s = ops.matte_loss_sums(a1, a4, a8, alphas, w1, w4, w8, valid.reshape(-1) if valid is not None else None)
    L = {}
    total = 0.0
    if cfg.loss_alpha_w > 0:
        r = s[:, 0] / (s[:, 5] + 1e-8)
        L.update(loss_rec_os1=r[0], loss_rec_os4=r[1], loss_rec_os8=r[2], loss_rec=r[0] * 2 + r[1] + r[2])
        total = total + L["loss_rec"] * cfg.loss_alpha_w
    if cfg.loss_alpha_lap_w > 0:
        # factor 3: the reference's LapLoss() is built for 3 channels and fed 1-channel images, so its [3,1,5,5] kernel
        # with groups=1 yields three identical channels and every level's weighted sum is counted three times
        lap = 3.0 * (s[:, 1:4] / (s[:, 5:8] + 1e-6)).sum(1)
        L.update(loss_lap_os1=lap[0], loss_lap_os4=lap[1], loss_lap_os8=lap[2], loss_lap=lap[0] * 2 + lap[1] + lap[2])
        total = total + L["loss_lap"] * cfg.loss_alpha_lap_w
    if cfg.loss_alpha_grad_w > 0:
        g = s[:, 4] / (s[:, 5] + 1e-6)
        L.update(loss_grad_os1=g[0], loss_grad_os4=g[1], loss_grad_os8=g[2], loss_grad=g[0] * 2 + g[1] + g[2])
        total = total + L["loss_grad"] * cfg.loss_alpha_grad_w
    L["total"] = total
    return L
