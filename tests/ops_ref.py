"""Reference (oracle-backed, CPU) implementations of the NATIVE ops' contracts.  Test-only: used (a) to check
the CUDA kernels against on the GPU box and (b) injected into `maggie_b200.ops` so the host-side model logic
can be exercised on the CPU container."""
import contextlib

import numpy as np
import torch
import torch.nn.functional as F

from maggie_b200 import ops
from oracle import unknown as U


def unknown_mask(alpha, widths, and_mask=None, alt=None, use_alt=None):
    if alt is not None and int(use_alt) != 0:
        alpha = alt
    out = U.compute_unknown(alpha.detach().cpu().float().numpy(), list(widths))
    if and_mask is not None:
        out = out * (and_mask.cpu().numpy() != 0)
    return torch.from_numpy(out.astype(np.uint8)).to(alpha.device)


def fuse_stage(src, finer, coarser, widths, and_mask):
    w = unknown_mask(src, widths, and_mask=and_mask)
    return torch.where(w != 0, finer.detach().float(), coarser.detach().float()), w


def _imap(coords, slots, H, W):
    m = np.full((slots, H, W), -1, np.int64)
    if len(coords):
        m[coords[:, 0], coords[:, 1], coords[:, 2]] = np.arange(len(coords))
    return m


def _lookup(imap, s, y, x):
    S, H, W = imap.shape
    ok = (y >= 0) & (y < H) & (x >= 0) & (x < W)
    out = np.full(len(s), -1, np.int64)
    out[ok] = imap[s[ok], y[ok], x[ok]]
    return out


def sites_tables_ref(roi):
    """numpy restatement of mg_sites_count + mg_sites_tables. roi uint8 [S,H,W] (numpy)."""
    S, H, W = roi.shape
    coords = [U.active_sites(roi)]
    shapes = [(H, W)]
    for _ in range(3):
        c, shp = U.downscale_sites(coords[-1], *shapes[-1])
        coords.append(c), shapes.append(shp)
    imaps = [_imap(c, S, *shp) for c, shp in zip(coords, shapes)]
    nbr, parent, child = [None] * 4, [None] * 4, [None] * 4
    for l in range(4):
        c = coords[l].astype(np.int64)
        s, y, x = c[:, 0], c[:, 1], c[:, 2]
        if l in (0, 2):
            nbr[l] = np.stack([_lookup(imaps[l], s, y + ky - 1, x + kx - 1) for ky in range(3) for kx in range(3)], 1)
        if l < 3:
            cols = []
            for ky in range(3):
                for kx in range(3):
                    ty, tx = y + 1 - ky, x + 1 - kx
                    r = _lookup(imaps[l + 1], s, ty // 2, tx // 2)
                    r[(ty % 2 != 0) | (tx % 2 != 0) | (ty < 0) | (tx < 0)] = -1
                    cols.append(r)
            parent[l] = np.stack(cols, 1)
        if l >= 1:
            child[l] = np.stack([_lookup(imaps[l - 1], s, 2 * y - 1 + ky, 2 * x - 1 + kx) for ky in range(3) for kx in range(3)], 1)
    return coords, nbr, parent, child, shapes


def build_sites(roi, status=None):
    dev = roi.device
    flags = None
    if status is not None:
        flags = [int(v) for v in status[4:].cpu()]
        if flags[ops.STATUS_EMPTY_MASK - 4]:
            raise ValueError("Mask is empty")
    coords, nbr, parent, child, shapes = sites_tables_ref(roi.cpu().numpy().astype(np.uint8))
    tt = lambda a: None if a is None else torch.from_numpy(
        np.ascontiguousarray(a).astype(np.int32).reshape(len(a), a.shape[1] if a.ndim > 1 else -1)).to(dev)
    return ops.SiteTables([len(c) for c in coords], [tt(c) for c in coords], [tt(a) for a in nbr],
                          [tt(a) for a in parent], [tt(a) for a in child], shapes, flags)


def mask_embed(image, masks, table, slot_ids, C=8, dtype=None):
    """Differentiable torch restatement of K1 (encoder/resnet.py:211-229); NCHW-shaped [B,C,H,W]."""
    B, _, H, W = image.shape
    slot_ids = slot_ids.tolist() if torch.is_tensor(slot_ids) else slot_ids
    ids = torch.tensor([s + 1 for s in slot_ids], device=image.device).view(1, -1, 1, 1)
    m = (masks * ids).long()
    on = (m > 0).float().unsqueeze(-1)
    emb = (table[m] * on).sum(1) / (on.sum(1) + 1e-6)
    out = torch.cat([image, emb.permute(0, 3, 1, 2), image.new_zeros(B, C - 6, H, W)], 1)
    return out.to(ops.COMPUTE_DTYPE if dtype in (None, torch.float16) else dtype).contiguous(memory_format=torch.channels_last)


def conv_bn_act(x, w, bn, training, *, stride=1, padding=1, dilation=1, act="relu", act_first=False,
                residual=None, transposed=False, res_up=False):
    """torch restatement of the conv+BN+act contract (cuDNN/ATen or CPU)."""
    import torch.nn.functional as F
    a = lambda t: F.relu(t) if act == "relu" else (F.leaky_relu(t, 0.2) if act == "lrelu" else t)
    w = w.to(x.dtype)
    cin_w = w.shape[0 if transposed else 1]
    if cin_w < x.shape[1]:  # input was channel-padded by mask_embed
        pad = x.shape[1] - cin_w
        w = F.pad(w, (0, 0, 0, 0, 0, pad)) if not transposed else F.pad(w, (0, 0, 0, 0, 0, 0, 0, pad))
    y = F.conv_transpose2d(x, w, stride=2, padding=1) if transposed else F.conv2d(x, w, stride=stride, padding=padding, dilation=dilation)
    if act_first:
        y = a(y)
    if bn is not None:
        y = ops.batch_norm(y, bn, training)
    if residual is not None:
        y = y + (F.interpolate(residual, scale_factor=2, mode="nearest") if res_up else residual)
    return y if act_first else a(y)


# ---- torch restatements of the sparse (rulebook) ops -------------------------------------------------------
def gather_conv(src, table, weight, bias=None):
    """Rulebook convolution: out[p] = sum_t W[t] . src[table[p,t]] (table entry -1 = no contribution).
    src [Ns,Cin]; table [No,T] int32; weight in spconv layout [Cout,kh,kw,Cin] with kh*kw == T.
    Covers SubMConv2d (nbr table) and SparseInverseConv2d (parent table).  torch reference."""
    No, T = table.shape
    Cout, Cin = weight.shape[0], weight.shape[-1]
    if No == 0:
        return src.new_zeros((0, Cout))
    padded = torch.cat([src, src.new_zeros((1, Cin))], dim=0)
    idx = torch.where(table < 0, src.shape[0], table.long()).reshape(-1)
    g = padded.index_select(0, idx).reshape(No, T * Cin)
    w = weight.reshape(Cout, T * Cin).to(src.dtype)
    out = g @ w.t()
    return out if bias is None else out + bias.to(out.dtype)


def pointwise_conv(src, weight, bias=None):
    """SubMConv2d with k=1 == per-site linear map. weight [Cout,1,1,Cin]. torch reference."""
    return ops.linear(src, weight.reshape(weight.shape[0], -1), bias)


def gather_dense(dense, coords, n_i):
    """dense NCHW-shaped channels-last [B,C,H,W] -> rows [N,C] at coords (frame = slot // n_i). torch reference."""
    nhwc = dense.permute(0, 2, 3, 1)
    c = coords.long()
    return nhwc[torch.div(c[:, 0], n_i, rounding_mode="floor"), c[:, 1], c[:, 2]]


def scatter_logits(vals, coords, slots, H, W):
    """fp32 logit map [slots,1,H,W]: -99 everywhere, value at active sites computed as ((v - 99) + 99) like the
    reference's dense()/-99/+=99 sequence (decoder/resnet_inst_matt_spconv.py:248-251). torch reference."""
    out = torch.full((slots, 1, H, W), -99.0, dtype=torch.float32, device=vals.device)
    c = coords.long()
    v = (vals.float().reshape(-1) - 99.0) + 99.0
    return out.index_put((c[:, 0], torch.zeros_like(c[:, 0]), c[:, 1], c[:, 2]), v)



def rows_conv(src, w, bias=None, *, table=None, table_t=None, mirror=False, bn=None, mode="plain", act=None, training=False):
    import torch.nn.functional as F
    co, ci = w.shape[0], w.shape[-1]
    y = gather_conv(src, table, w.reshape(co, -1, 1, ci), bias) if table is not None else pointwise_conv(src, w.reshape(co, 1, 1, ci), bias)
    if mode == "plain":
        return y
    a = lambda t: F.relu(t) if act == "relu" else (F.leaky_relu(t, 0.2) if act == "lrelu" else t)
    if mode == "act_bn":
        y = F.relu(y)
    if y.shape[0]:
        y = ops.batch_norm(y.float(), bn, training).to(y.dtype)
    return a(y) if mode == "bn_act" else y


def rows_head(src, w, bias, nbr, coords, slots, H, W):
    y = gather_conv(src, nbr, w.reshape(1, 9, 1, w.shape[-1]), bias)
    return scatter_logits(y, coords, slots, H, W)


# ---- torch restatement of the attention core (K6) ----------------------------------------------------------
def attention(q, k, v, key_padding=None, need_stat=None):
    """Single-head attention, batch-first: q [B,L,E], k/v [B,S,E] (already projected).  Softmax in fp32.
    key_padding [B,S] bool (True = ignore).  need_stat: optional [B,L,S] bool guidance mask; if given also
    returns stat[b,l] = sum_s guidance * A (the only thing the attention-max loss needs,
    module/instance_matte_decoder.py:101-109).  torch reference."""
    s = torch.bmm(q, k.transpose(1, 2)).float() * (q.shape[-1] ** -0.5)
    if key_padding is not None:
        s = s.masked_fill(key_padding[:, None, :], float("-inf"))
    a = torch.softmax(s, dim=-1)
    o = torch.bmm(a.to(v.dtype), v)
    if need_stat is not None:
        return o, (a * need_stat).sum(-1)
    return o, None



# ---- torch restatement of the K12 partial sums (loss.py:67-191 semantics) -------------------------------------
_G5 = torch.tensor([1.0, 4.0, 6.0, 4.0, 1.0])
_GAUSS2D = (_G5[:, None] * _G5[None, :]) / 256.0
_SOBEL = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]]) / 8.0


def _blur(x, k):
    return F.conv2d(F.pad(x, (2, 2, 2, 2), mode="reflect"), k)


def _pyramid(x, levels=3):
    k = _GAUSS2D.to(x.device, x.dtype)[None, None]
    out = []
    for _ in range(levels):
        down = _blur(x, k)[:, :, ::2, ::2]
        up = x.new_zeros(x.shape)
        up[:, :, ::2, ::2] = down
        out.append(x - _blur(up, 4.0 * k))
        x = down
    return out


def _sobel_mag(x, eps=1e-6):
    xp = F.pad(x, (1, 1, 1, 1), mode="replicate")
    kx = _SOBEL.to(x.device, x.dtype)
    gx, gy = F.conv2d(xp, kx[None, None]), F.conv2d(xp, kx.t()[None, None])
    return torch.sqrt(gx * gx + gy * gy + eps)


def matte_loss_sums(a1, a4, a8, target, w1, w4, w8, plane_scale=None):
    if plane_scale is not None:
        ps = plane_scale.reshape(a1.shape[:-2] + (1, 1)).to(a1.dtype)
        a1, a4, a8 = a1 * ps, a4 * ps, a8 * ps
    h, w = a1.shape[-2:]
    t = target.reshape(-1, 1, h, w).float()
    rows = []
    for p, wt in ((a1, w1), (a4, w4), (a8, w8)):
        p, wt = p.reshape(-1, 1, h, w).float(), wt.reshape(-1, 1, h, w).float()
        q = [(p * wt - t * wt).abs().sum()]
        pyr, wl, ws = _pyramid(p - t), wt, []
        for i in range(3):
            q.append((pyr[i].abs() * wl).sum())
            ws.append(wl.sum())
            wl = wl[:, :, ::2, ::2]
        q.append((_sobel_mag(p * wt) - _sobel_mag(t * wt)).abs().sum())
        rows.append(torch.stack(q + ws))
    return torch.stack(rows)


def conv_bias(x, w, bias=None, *, padding=1):
    return F.conv2d(x, w.to(x.dtype), None if bias is None else bias.to(x.dtype), padding=padding)


@contextlib.contextmanager
def injected(dtype=torch.float32):
    """Swap the native ops for the references above (CPU container only)."""
    names = ("unknown_mask", "build_sites", "mask_embed", "conv_bn_act", "rows_conv", "rows_head", "gather_dense",
             "matte_loss_sums", "attention", "conv_bias", "prepare_weights", "COMPUTE_DTYPE", "fuse_stage")
    saved = {n: getattr(ops, n) for n in names}
    ops.unknown_mask, ops.build_sites, ops.mask_embed, ops.conv_bn_act, ops.COMPUTE_DTYPE = \
        unknown_mask, build_sites, mask_embed, conv_bn_act, dtype
    ops.fuse_stage = fuse_stage
    ops.rows_conv, ops.rows_head, ops.gather_dense = rows_conv, rows_head, gather_dense
    ops.matte_loss_sums, ops.attention, ops.conv_bias = matte_loss_sums, attention, conv_bias
    ops.prepare_weights = lambda bank: None  # layers then use the per-layer torch composition (ops.spectral_weight)
    try:
        yield
    finally:
        for n, v in saved.items():
            setattr(ops, n, v)
