"""Native sparse-refinement ops (K9) - python side: rulebook convolutions on [N, C] fp16 site rows with
autograd, BatchNorm1d through the K3 kernels, dense<->sparse gathers.  Weight tensors keep the spconv layout
[Cout, kh, kw, Cin] (Linear: [out, in])."""
import ctypes
from ctypes import c_int32, c_void_p

import torch

from . import _lib, dense, packs
from .dense import ACT, bn_finalize, exchange_sums, new_stats, sync_group, zeros_f32

_need_cuda, _ptr, _stream = _lib.need_cuda, _lib.tensor_ptr, _lib.stream_ptr


class SparseConvDesc(ctypes.Structure):
    """Mirror of `mg_sparse_conv_desc`."""
    _fields_ = [
        ("src", c_void_p), ("src_stride", c_int32),
        ("table", c_void_p), ("T", c_int32), ("No", c_int32), ("Cin", c_int32), ("Cout", c_int32),
        ("w", c_void_p), ("bias", c_void_p),
        ("out", c_void_p), ("out_stride", c_int32), ("c_off", c_int32),
        ("stats", c_void_p),
        ("map", c_void_p), ("coords", c_void_p), ("mapH", c_int32), ("mapW", c_int32),
        ("pre_act", c_int32),
    ]


def _pack_rows(w2d):
    """[Cout, K] fp32 -> fp16 [ceil16(Cout), K] (zero rows appended)."""
    co = w2d.shape[0]
    p = w2d.to(torch.float16)
    if co % 16:
        p = torch.cat([p, p.new_zeros(16 - co % 16, p.shape[1])], 0)
    return p.contiguous()


def pack_fwd(w):
    """spconv [Cout,kh,kw,Cin] (or Linear [out,in]) -> [ceil16(Cout), T*Cin]."""
    return _pack_rows(w.reshape(w.shape[0], -1))


def pack_bwd(w, mirror):
    """-> [Cin, T*Cout_pad] for the data gradient; `mirror` reverses the tap order (SubM 3x3: nbr table reuse)."""
    co, ci = w.shape[0], w.shape[-1]
    w3 = w.reshape(co, -1, ci)                       # [Cout, T, Cin]
    if mirror:
        w3 = w3.flip(1)
    cop = max(32, (co + 31) // 32 * 32)              # the gradient rows are the K dimension: multiple of 32
    wt = w3.permute(2, 1, 0)                          # [Cin, T, Cout]
    if cop != co:
        wt = torch.cat([wt, wt.new_zeros(ci, wt.shape[1], cop - co)], 2)
    return _pack_rows(wt.reshape(ci, -1)), cop


def _packs_for(w):
    """(pack set, entry) when the grouped packs of the running stage cover this weight (maggie_b200/packs.py)."""
    ps = packs.active()
    if ps is not None:
        e = ps.lookup(w)
        if e is not None:
            return ps, e
    return None, None


def _fwd_pack(w):
    """-> (forward pack, handle for the backward pack: the gathered buffer of THIS forward + the layer's entry)."""
    ps, e = _packs_for(w)
    return (ps.pack_fwd(e), (ps.bwd, e)) if ps is not None else (pack_fwd(w), None)


def _bwd_pack(w, mirror, handle):
    if handle is not None and handle[0] is not None:
        buf, e = handle
        o, shape = e["bwd", bool(mirror)]
        return buf[o:o + shape[0] * shape[1]].view(shape), e["cop"]
    return pack_bwd(w, mirror)


def sparse_conv_launch(src, wp, T, cin, cout, *, table=None, bias=None, out=None, out_stride=None, c_off=0, stats=None,
                       pre_act=0, head=None, n_out=None):
    """Raw launch.  src fp16 [Ns, >=cin] (row stride = src.stride(0)); wp packed weights; table int32 [No, T] or None.
    head = (map fp32 [slots,H,W], coords int32 [No,3]) switches to the logit-map epilogue."""
    No = n_out if n_out is not None else (table.shape[0] if table is not None else src.shape[0])
    if _dense_rows(src, table, T, cin, cout, No, head, out):
        # a plain GEMM over contiguous rows (the pixel-side attention projections: 32768 x 128 -> 128): the rows are an
        # NHWC image [1, No / W, W, cin] and the 1x1 case of the dense conv kernel (TMA tiles instead of per-row
        # cp.async gathers) is 2x faster than the rulebook kernel (7 vs 16 us)
        from . import dense
        W = 64 if No % 64 == 0 else 16
        if out is None:
            out = torch.empty((No, cout), dtype=torch.float16, device=src.device)
        dense.conv_launch(src.view(1, No // W, W, cin), wp, [(0, 0, 0)], out=out.view(1, No // W, W, out.shape[1]), c_off=c_off,
                          stats=stats, bias=bias, pre_act="relu" if pre_act == 1 else None)
        return out
    d = SparseConvDesc()
    d.src, d.src_stride = src.data_ptr(), src.stride(0)
    d.table = table.data_ptr() if table is not None else None
    d.T, d.No, d.Cin, d.Cout = T, No, cin, cout
    d.w = wp.data_ptr()
    d.bias = bias.data_ptr() if bias is not None else None
    if head is not None:
        mp, coords = head
        d.map, d.coords, d.mapH, d.mapW = mp.data_ptr(), coords.data_ptr(), mp.shape[-2], mp.shape[-1]
        d.out, d.out_stride, d.c_off = None, 0, 0
    else:
        if out is None:
            out = torch.empty((No, cout), dtype=torch.float16, device=src.device)
        d.out, d.out_stride, d.c_off = out.data_ptr(), out.stride(0), c_off
        d.map = d.coords = None
    d.stats = stats.data_ptr() if stats is not None else None
    d.pre_act = pre_act
    _lib.check(_lib.lib().mg_sparse_conv(ctypes.byref(d), _stream()), "mg_sparse_conv")
    return out


DENSE_ROWS = __import__("os").environ.get("MAGGIE_B200_NO_DENSE_ROWS", "0") != "1"


def _dense_rows(src, table, T, cin, cout, No, head, out):
    """True when a rows launch is a plain dense GEMM the conv kernels take (see sparse_conv_launch)."""
    return (DENSE_ROWS and table is None and T == 1 and head is None and No == src.shape[0] and No >= 8192 and No % 16 == 0
            and cin >= 128 and cin % 16 == 0 and cout % 16 == 0 and src.dim() == 2 and src.shape[1] == cin and src.is_contiguous()
            and (out is None or (out.dim() == 2 and out.is_contiguous())))


def _f16rows(t):
    if t.dtype != torch.float16 or t.stride(-1) != 1 or t.stride(0) % 8:
        t = t.to(torch.float16).contiguous()
    return t


def _wgrad(dout, cout, src, cin, table, T):
    from . import dense
    dw = dense.zeros_f32(cout * T * cin, src.device).view(cout, T * cin)
    No = dout.shape[0]
    if (_dense_rows(src, table, T, cin, cout, No, None, None) and dout.dim() == 2 and dout.shape[1] == cout
            and dout.is_contiguous()):
        W = 64 if No % 64 == 0 else 16   # the dense weight-gradient kernel (K4) on the rows viewed as an image
        dense.wgrad_launch(dout.view(1, No // W, W, cout), src.view(1, No // W, W, cin), [(0, 0, 0)], dw, grid_hw=(No // W, W))
        return dw
    _lib.check(_lib.lib().mg_sparse_wgrad(_ptr(dout), dout.stride(0), cout, _ptr(src), src.stride(0), cin,
                                         _ptr(table) if table is not None else None, T, No, _ptr(dw), _stream()),
               "mg_sparse_wgrad")
    return dw


class _RowsConv(torch.autograd.Function):
    """Rulebook conv (+ optional BatchNorm1d with batch statistics and activation), all native.
    mode: 'plain' (conv + bias), 'bn_act' (conv -> BN -> act), 'act_bn' (conv + bias -> ReLU -> BN)."""

    @staticmethod
    def forward(ctx, src, w, bias, gamma, beta, table, table_t, mirror, bn, mode, act, training):
        src = _f16rows(src)
        co, ci = w.shape[0], w.shape[-1]
        T = w.numel() // (co * ci)
        wd = w.detach()
        wp, ctx.packs = _fwd_pack(wd)
        b32 = bias.detach().float().contiguous() if bias is not None else None
        No = table.shape[0] if table is not None else src.shape[0]
        ctx.cfg = (mode, act, T, co, ci, mirror, tuple(w.shape), bias is not None, training)
        if mode == "plain":
            y = sparse_conv_launch(src, wp, T, ci, co, table=table, bias=b32)
            ctx.save_for_backward(src, wd, table, table_t)
            return y
        if not training:  # eval: running statistics -> one affine pass after the conv
            r = sparse_conv_launch(src, wp, T, ci, co, table=table, bias=b32, pre_act=1 if mode == "act_bn" else 0)
            scale, shift, _, _ = bn_finalize(None, 1, bn, False)
            y = torch.empty_like(r)
            _lib.check(_lib.lib().mg_bn_apply(_ptr(r), _ptr(scale), _ptr(shift), None, 0, _ptr(y), No, 1, 1, co,
                                              0 if mode == "act_bn" else ACT[act], _stream()), "mg_bn_apply")
            return y
        group = sync_group(bn)   # SyncBatchNorm-equivalent: statistics over the active sites of all ranks (dense.py)
        stats = new_stats(co, src.device)
        r = sparse_conv_launch(src, wp, T, ci, co, table=table, bias=b32, stats=stats, pre_act=1 if mode == "act_bn" else 0)
        y = torch.empty_like(r)
        if group is None and dense.FUSED_BN_APPLY and No > 0 and co <= 512:
            # local statistics: finalize + apply in one launch (dense.py, _ConvBNAct)
            dense.bump_counter(bn.num_batches_tracked)
            out4 = torch.empty((4, co), dtype=torch.float32, device=src.device)
            _lib.check(_lib.lib().mg_bn_train_apply(_ptr(stats), float(No), _ptr(bn.weight), _ptr(bn.bias), _ptr(bn.running_mean),
                                                    _ptr(bn.running_var), float(bn.momentum), float(bn.eps), _ptr(out4), _ptr(r),
                                                    None, 0, _ptr(y), No, 1, 1, co, 0 if mode == "act_bn" else ACT[act],
                                                    _stream()), "mg_bn_train_apply")
            mean, invstd = out4[2], out4[3]
            ctx.sync = None
        else:
            scale, shift, mean, invstd, *cnt = bn_finalize(stats, max(No, 1) if group is None else No, bn, True, group)
            ctx.sync = (group, cnt[0]) if group is not None else None
            _lib.check(_lib.lib().mg_bn_apply(_ptr(r), _ptr(scale), _ptr(shift), None, 0, _ptr(y), No, 1, 1, co,
                                              0 if mode == "act_bn" else ACT[act], _stream()), "mg_bn_apply")
        ctx.save_for_backward(src, wd, table, table_t, r, y, mean, invstd, gamma.detach())
        ctx.sums = zeros_f32(2 * co, src.device).view(2, co)  # zeroed now (scope pool), filled by the backward
        return y

    @staticmethod
    def backward(ctx, gy):
        mode, act, T, co, ci, mirror, w_shape, has_bias, training = ctx.cfg
        L = _lib.lib()
        if mode == "plain":
            src, w, table, table_t = ctx.saved_tensors
            dr = _f16rows(gy)
            dgamma = dbeta = None
        else:
            src, w, table, table_t, r, y, mean, invstd, gamma = ctx.saved_tensors
            dy = _f16rows(gy)
            No = r.shape[0]
            sums, ctx.sums = ctx.sums, None
            if sums is None:
                sums = zeros_f32(2 * co, r.device).view(2, co)
            a_post = 0 if mode == "act_bn" else ACT[act]
            _lib.check(L.mg_bn_bwd_reduce(_ptr(dy), _ptr(y), _ptr(r), _ptr(mean), _ptr(invstd), _ptr(sums), No, 1, 1, co,
                                          a_post, _stream()), "mg_bn_bwd_reduce")
            dr = torch.empty_like(r)
            gsums, count_dev = sums, None
            if ctx.sync is not None:
                gsums, count_dev = exchange_sums(sums, 1, co, None, ctx.sync[0]), ctx.sync[1]
            _lib.check(L.mg_bn_bwd_apply(_ptr(dy), _ptr(y), _ptr(r), _ptr(mean), _ptr(invstd), _ptr(gamma), _ptr(gsums),
                                         _ptr(dr), None, No, 1, 1, co, a_post, 1 if mode == "act_bn" else 0, _ptr(count_dev),
                                         _stream()), "mg_bn_bwd_apply")
            dgamma, dbeta = sums[1], sums[0]
        dsrc = dw = dbias = None
        if ctx.needs_input_grad[0]:
            wt, cop = _bwd_pack(w, mirror, ctx.packs)
            g = dr
            if cop != co:  # pad the gradient rows to the K granularity
                g = torch.zeros((dr.shape[0], cop), dtype=torch.float16, device=dr.device)
                g[:, :co] = dr
            tt = table_t if table is not None else None
            n_src = src.shape[0]
            dsrc = sparse_conv_launch(g, wt, T, cop, ci, table=tt, n_out=n_src)
        if ctx.needs_input_grad[1]:
            dw = _wgrad(dr, co, src, ci, table, T).view(w_shape)
        if has_bias and ctx.needs_input_grad[2]:
            from . import ops
            dbias = ops.col_sum(dr)
        return dsrc, dw, dbias, dgamma, dbeta, None, None, None, None, None, None, None


def rows_conv(src, w, bias=None, *, table=None, table_t=None, mirror=False, bn=None, mode="plain", act=None, training=False):
    """src [Ns, Cin] rows -> [No, Cout] rows.  SubM 3x3: table = table_t = nbr, mirror=True.  Inverse conv:
    table = parent[l], table_t = child[l+1].  1x1 / Linear: no tables."""
    _need_cuda(src, w)
    gamma = bn.weight if bn is not None else None
    beta = bn.bias if bn is not None else None
    return _RowsConv.apply(src, w, bias, gamma, beta, table, table_t, mirror, bn, mode, act, training)


class _Head(torch.autograd.Function):
    """SubM 3x3 Cin -> 1 (+bias) written straight into the fp32 logit map (-99 where inactive)."""

    @staticmethod
    def forward(ctx, src, w, bias, nbr, coords, slots, H, W):
        src = _f16rows(src)
        ci = w.shape[-1]
        wd = w.detach()
        mp = torch.full((slots, 1, H, W), -99.0, dtype=torch.float32, device=src.device)
        wp, ctx.packs = _fwd_pack(wd)
        sparse_conv_launch(src, wp, 9, ci, 1, table=nbr, bias=bias.detach().float().contiguous(), head=(mp, coords))
        ctx.save_for_backward(src, wd, nbr, coords)
        return mp

    @staticmethod
    def backward(ctx, gmap):
        src, w, nbr, coords = ctx.saved_tensors
        ci = w.shape[-1]
        c = coords.long()
        g = torch.zeros((coords.shape[0], 32), dtype=torch.float16, device=src.device)
        g[:, 0] = gmap[c[:, 0], 0, c[:, 1], c[:, 2]].to(torch.float16)
        wt, cop = _bwd_pack(w, True, ctx.packs)
        dsrc = sparse_conv_launch(g, wt, 9, cop, ci, table=nbr, n_out=src.shape[0]) if ctx.needs_input_grad[0] else None
        dw = _wgrad(g, 8, src, ci, nbr, 9)[:1].view(w.shape) if ctx.needs_input_grad[1] else None
        dbias = g[:, 0].float().sum().reshape(1) if ctx.needs_input_grad[2] else None
        return dsrc, dw, dbias, None, None, None, None, None


def rows_head(src, w, bias, nbr, coords, slots, H, W):
    _need_cuda(src, w)
    return _Head.apply(src, w, bias, nbr, coords, slots, H, W)


class _GatherDense(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dense, coords, n_i):
        dn = dense.permute(0, 2, 3, 1)
        assert dn.is_contiguous() and dn.dtype == torch.float16
        B, H, W, C = dn.shape
        n = coords.shape[0]
        out = torch.empty((n, C), dtype=torch.float16, device=dense.device)
        _lib.check(_lib.lib().mg_gather_rows(_ptr(dn), _ptr(coords), n, n_i, H, W, C, _ptr(out), C, 0, _stream()),
                   "mg_gather_rows")
        ctx.save_for_backward(coords)
        ctx.meta = (n_i, (B, H, W, C))
        return out

    @staticmethod
    def backward(ctx, g):
        (coords,) = ctx.saved_tensors
        n_i, (B, H, W, C) = ctx.meta
        g = _f16rows(g)
        dd = torch.zeros((B, H, W, C), dtype=torch.float16, device=g.device)
        _lib.check(_lib.lib().mg_scatter_rows_add(_ptr(g), g.stride(0), 0, _ptr(coords), coords.shape[0], n_i, H, W, C,
                                                 _ptr(dd), _stream()), "mg_scatter_rows_add")
        return dd.permute(0, 3, 1, 2), None, None


def gather_dense(dense, coords, n_i):
    """dense NCHW-shaped channels-last fp16 [B,C,H,W] -> rows [N,C] at coords (frame = slot // n_i)."""
    _need_cuda(dense, coords)
    if dense.dtype != torch.float16 or not dense.permute(0, 2, 3, 1).is_contiguous():
        dense = dense.to(torch.float16).contiguous(memory_format=torch.channels_last)
    return _GatherDense.apply(dense, coords, n_i)
