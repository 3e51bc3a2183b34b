"""numpy restatement of the uncertainty ("unknown" / incoherence) mask and the active-site list.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows maggie/utils/utils.py:27-55 (`Kernels`,
`compute_unknown`) and maggie/network/decoder/resnet_inst_matt_spconv.py:203-213 (`torch.nonzero`).
`cv2.getStructuringElement(MORPH_ELLIPSE)` and `cv2.dilate` are OpenCV (4.13 in this image, absent from
/root/reference): the published algorithm is restated here and pinned against cv2 itself in
tests/test_oracle_unknown.py for every kernel size 1..29 the reference can draw.
"""
import numpy as np

LOWER, UPPER = np.float32(1.0 / 255.0), np.float32(254.0 / 255.0)  # utils.py:28 defaults


def ellipse_spans(k):
    """Row spans [j1, j2) of cv2.getStructuringElement(MORPH_ELLIPSE, (k, k)); anchor is (k//2, k//2).
    OpenCV: r = k/2, c = k/2 (integer), dx = saturate_cast<int>(c*sqrt((r*r - dy*dy)/r^2)) (round-half-even)."""
    r = c = k // 2
    inv_r2 = 1.0 / (r * r) if r else 0.0
    spans = []
    for i in range(k):
        dy = i - r
        if abs(dy) <= r:
            dx = int(np.rint(c * np.sqrt((r * r - dy * dy) * inv_r2)))
            spans.append((max(c - dx, 0), min(c + dx + 1, k)))
        else:
            spans.append((0, 0))
    return spans


def ellipse_kernel(k):
    m = np.zeros((k, k), np.uint8)
    for i, (j1, j2) in enumerate(ellipse_spans(k)):
        m[i, j1:j2] = 1
    return m


def dilate(u, k):
    """Binary dilation of uint8 [H, W] by ELLIPSE(k), anchor k//2, border ignored (cv2.dilate default)."""
    H, W = u.shape
    a = k // 2
    pad = k
    cs = np.zeros((H, W + 2 * pad + 1), np.int32)
    cs[:, pad + 1:pad + 1 + W] = u
    cs = np.cumsum(cs, axis=1)
    out = np.zeros((H, W), bool)
    xs = np.arange(W) + pad
    for i, (j1, j2) in enumerate(ellipse_spans(k)):
        if j2 <= j1:
            continue
        # out[y, x] |= any(u[y + i - a, x + j - a] for j in [j1, j2))
        row_any = (cs[:, xs + (j2 - 1 - a) + 1] - cs[:, xs + (j1 - a)]) > 0
        dy = i - a
        ys0, ys1 = max(0, -dy), min(H, H - dy)
        if ys1 > ys0:
            out[ys0:ys1] |= row_any[ys0 + dy:ys1 + dy]
    return out.astype(np.uint8)


def compute_unknown(alpha, widths):
    """alpha: float array [..., H, W]; widths: one ellipse size per [H, W] slice (flattened order).
    Returns uint8 array of the same shape (utils.py:28-55)."""
    a = np.asarray(alpha, np.float32)
    H, W = a.shape[-2:]
    u = ((a > LOWER) & (a < UPPER)).astype(np.uint8).reshape(-1, H, W)
    widths = np.broadcast_to(np.asarray(widths, np.int64), (u.shape[0],))
    out = np.stack([dilate(u[n], int(widths[n])) for n in range(u.shape[0])]) if u.shape[0] else u
    return out.reshape(a.shape)


def active_sites(roi):
    """roi: uint8 [slots, H, W] -> int32 [N, 3] rows (slot, y, x) in lexicographic order (torch.nonzero)."""
    return np.argwhere(np.asarray(roi) > 0).astype(np.int32)


def downscale_sites(sites, H, W):
    """Index set produced by SparseConv2d(k=3, s=2, p=1) on `sites` (spconv semantics, see
    oracle/spconv_torch.py): q active iff any active input in rows/cols 2q-1..2q+1. Sorted rows."""
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    keys = []
    s, y, x = sites[:, 0].astype(np.int64), sites[:, 1].astype(np.int64), sites[:, 2].astype(np.int64)
    for ky in range(3):
        for kx in range(3):
            ty, tx = y + 1 - ky, x + 1 - kx
            ok = (ty % 2 == 0) & (tx % 2 == 0)
            qy, qx = ty // 2, tx // 2
            ok &= (qy >= 0) & (qy < Ho) & (qx >= 0) & (qx < Wo)
            keys.append((s[ok] * Ho + qy[ok]) * Wo + qx[ok])
    uniq = np.unique(np.concatenate(keys)) if keys else np.zeros(0, np.int64)
    return np.stack([uniq // (Ho * Wo), (uniq // Wo) % Ho, uniq % Wo], axis=1).astype(np.int32), (Ho, Wo)
