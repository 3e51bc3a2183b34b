"""Training losses evaluated inside forward(): weighted L1, 3-level Laplacian pyramid, Sobel gradient.

Reference: arch/maggie.py:237-368 (regression_loss, compute_loss), loss.py:67-191 (GradientLoss, LapLoss).
INTERIM: composed from torch ops in fp32 (a fused stencil-reduction kernel is the SURVEY §8f rank-1 "next" row).
"""
import torch
import torch.nn.functional as F

_GAUSS = torch.tensor([1.0, 4.0, 6.0, 4.0, 1.0])
_GAUSS2D = (_GAUSS[:, None] * _GAUSS[None, :]) / 256.0
_SOBEL = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]]) / 8.0


def weighted_l1(pred, target, weight):
    return (pred * weight - target * weight).abs().sum() / (weight.sum() + 1e-8)


def _dwconv(xp, k):
    """Single-channel stencil over N images run as ONE depthwise conv over N channels (direct kernels) instead of
    an N-batch 1->1 channel implicit GEMM."""
    n = xp.shape[0]
    return F.conv2d(xp.transpose(0, 1), k.expand(n, 1, *k.shape[-2:]), groups=n).transpose(0, 1)


def _blur(x, k):
    return _dwconv(F.pad(x, (2, 2, 2, 2), mode="reflect"), k)


def _pyramid(x, levels):
    k = _GAUSS2D.to(x.device, x.dtype)[None, None]
    out = []
    for _ in range(levels):
        down = _blur(x, k)[:, :, ::2, ::2]
        up = x.new_zeros(x.shape)
        up[:, :, ::2, ::2] = down
        out.append(x - _blur(up, 4.0 * k))
        x = down
    return out


def lap_loss(pred, target, weight, levels=3):
    """pred/target/weight [N,1,H,W].  The reference's LapLoss() is built for 3 channels and fed 1-channel images,
    which triples every level's weighted sum (loss.py:170-173 with groups=1) - reproduced by the factor 3."""
    pp, pt = _pyramid(pred, levels), _pyramid(target, levels)
    total = 0.0
    for i in range(levels):
        total = total + 3.0 * ((pp[i] - pt[i]).abs() * weight).sum() / (weight.sum() + 1e-6)
        weight = weight[:, :, ::2, ::2]
    return total


def _sobel_mag(x, eps=1e-6):
    n, c, h, w = x.shape
    xp = F.pad(x.reshape(n * c, 1, h, w), (1, 1, 1, 1), mode="replicate")
    kx = _SOBEL.to(x.device, x.dtype)
    gx, gy = _dwconv(xp, kx[None, None]), _dwconv(xp, kx.t()[None, None])
    return torch.sqrt(gx * gx + gy * gy + eps).reshape(n, c, h, w)


def grad_loss(pred, target, weight, eps=1e-6):
    return (_sobel_mag(pred * weight) - _sobel_mag(target * weight)).abs().sum() / (weight.sum() + eps)


def compute_loss(pred, w4, w1, alphas, cfg):
    """Image-model loss dictionary (dtSSD handled by the video subclass).  The three scales are evaluated as one
    batch, and - the Laplacian pyramid being linear - the pyramid is built once on (pred - target)."""
    a1, a4, a8 = pred["alpha_os1"], pred["alpha_os4"], pred["alpha_os8"]
    w8 = (alphas.sum((2, 3), keepdim=True) > 0).to(a8.dtype).expand_as(a8)
    if cfg.loss_reweight_os8:
        lo, hi = 1.0 / 255.0, 254.0 / 255.0
        unk = ((alphas <= hi) & (alphas >= lo)) | ((a8 <= hi) & (a8 >= lo))
        w8 = unk.to(a8.dtype) + w8
    h, w = a8.shape[-2:]
    P = torch.stack([a1, a4, a8]).reshape(3, -1, 1, h, w)                       # [3, S, 1, h, w]
    Wt = torch.stack([w1.to(a8.dtype), w4.to(a8.dtype), w8]).reshape(3, -1, 1, h, w)
    T = alphas.reshape(1, -1, 1, h, w)
    n = P.shape[1]
    per_scale = lambda z: z.reshape(3, -1).sum(1)
    wsum = per_scale(Wt)
    L = {}
    total = 0.0
    if cfg.loss_alpha_w > 0:
        r = per_scale((P * Wt - T * Wt).abs()) / (wsum + 1e-8)
        L.update(loss_rec_os1=r[0], loss_rec_os4=r[1], loss_rec_os8=r[2], loss_rec=r[0] * 2 + r[1] + r[2])
        total = total + L["loss_rec"] * cfg.loss_alpha_w
    if cfg.loss_alpha_lap_w > 0:
        pyr = _pyramid((P - T).reshape(3 * n, 1, h, w), 3)
        wl, lap = Wt.reshape(3 * n, 1, h, w), 0.0
        for i in range(3):
            # factor 3: the reference's LapLoss() is built for 3 channels and fed 1-channel images (see lap_loss)
            lap = lap + 3.0 * per_scale(pyr[i].abs() * wl) / (per_scale(wl) + 1e-6)
            wl = wl[:, :, ::2, ::2]
        L.update(loss_lap_os1=lap[0], loss_lap_os4=lap[1], loss_lap_os8=lap[2], loss_lap=lap[0] * 2 + lap[1] + lap[2])
        total = total + L["loss_lap"] * cfg.loss_alpha_lap_w
    if cfg.loss_alpha_grad_w > 0:
        sp = _sobel_mag((P * Wt).reshape(3 * n, 1, h, w))
        st = _sobel_mag((T * Wt).reshape(3 * n, 1, h, w))
        g = per_scale((sp - st).abs()) / (wsum + 1e-6)
        L.update(loss_grad_os1=g[0], loss_grad_os4=g[1], loss_grad_os8=g[2], loss_grad=g[0] * 2 + g[1] + g[2])
        total = total + L["loss_grad"] * cfg.loss_alpha_grad_w
    L["total"] = total
    return L
