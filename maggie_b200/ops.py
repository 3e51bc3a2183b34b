"""Operator layer between `maggie_b200.network` and the C ABI (include/maggie_b200.h).

Two kinds of ops live here:
  * NATIVE ops - `torch.autograd.Function`s / plain functions that hand raw device pointers to
    libmaggie_b200.so (hand-written sm_100a kernels).  No fallback: they raise if the tensors are not CUDA
    tensors or the library is missing.
  * INTERIM ops - the parts of the path whose kernels are not written yet are composed from torch CUDA ops
    (cuDNN/cuBLAS/ATen).  They are listed in DESIGN.md ("native coverage") and are replaced one by one; the
    model code above this layer does not change when that happens.
  * ROUTED ops (`layer_norm`, `token_logits`, `upsample_tanh`, `col_sum`) - a native kernel for the shapes the path
    uses and the equivalent torch composition for anything else (other shapes on the GPU).  The composition is also
    what runs when the CPU test-suite drives the model's host logic with CPU tensors (`tests/test_host_model.py`,
    where `tests/ops_ref.py` stands in for the native-only ops); a model on a GPU never takes it for the path's shapes,
    and a CUDA tensor with the library missing raises (`_lib.lib()`), it does not fall back.
Activations are fp16, channels-last in memory (NHWC) and NCHW-shaped for torch; statistics, logits and
alphas are fp32.
"""
import ctypes
from collections import namedtuple

import torch
import torch.nn.functional as F

from . import _lib

COMPUTE_DTYPE = torch.float16


_stream, _ptr, _need_cuda = _lib.stream_ptr, _lib.tensor_ptr, _lib.need_cuda


# =============================================================================================== native: K8a
_WIDTHS_CACHE = {}


def _widths_tensor(widths, device):
    """Device int32 copy of the per-slice ellipse sizes; constant (eval) width lists are cached, random ones (training)
    go through a pinned staging ring so that the copy is asynchronous."""
    key = (widths, device)
    t = _WIDTHS_CACHE.get(key)
    if t is None:
        if len(set(widths)) == 1 or len(widths) <= 16:
            t = torch.tensor(widths, dtype=torch.int32, device=device)
            if len(_WIDTHS_CACHE) > 256:
                _WIDTHS_CACHE.clear()
            _WIDTHS_CACHE[key] = t
        elif len(widths) <= _IntStager.CAP and torch.device(device).type == "cuda":
            st = _STAGERS.get(str(device))
            if st is None:
                st = _STAGERS[str(device)] = _IntStager(device)
            t = st.put(widths)
        else:
            t = torch.tensor(widths, dtype=torch.int32, device=device)
    return t


_INDEX_CACHE = {}


def index_tensor(idx, device):
    """Cached device int64 copy of a short python index list.  Indexing a CUDA tensor with a python list makes torch
    build the index tensor with a synchronous host-to-device copy, which drains the launch queue; the cache pays that
    once per distinct list (slot subsets: at most C(10, n) of them)."""
    key = (tuple(int(i) for i in idx), str(device))
    t = _INDEX_CACHE.get(key)
    if t is None:
        if len(_INDEX_CACHE) > 4096:
            _INDEX_CACHE.clear()
        t = _INDEX_CACHE[key] = torch.tensor(key[0], dtype=torch.long, device=device)
    return t


def take(t, dim, idx):
    """t.index_select(dim, idx) for a python index list, without the host synchronisation of `t[:, idx]`."""
    return t.index_select(dim, index_tensor(idx, t.device))


def put(out, dim, idx, src):
    """out.index_copy_(dim, idx, src) for a python index list (no host synchronisation)."""
    return out.index_copy_(dim, index_tensor(idx, out.device), src)


class _IntStager:
    """Ring of pinned host buffers + device buffers for small per-call int32 lists that change every call (the random
    ellipse sizes of training): `torch.tensor(list, device=cuda)` copies from pageable memory, i.e. synchronously."""
    SLOTS, CAP = 32, 1024

    def __init__(self, device):
        self.host = torch.empty((self.SLOTS, self.CAP), dtype=torch.int32).pin_memory()
        self.dev = torch.empty((self.SLOTS, self.CAP), dtype=torch.int32, device=device)
        self.events = [None] * self.SLOTS
        self.turn = 0

    def put(self, values):
        n = len(values)
        i = self.turn % self.SLOTS
        self.turn += 1
        if self.events[i] is not None:
            self.events[i].synchronize()       # long done unless the host is > SLOTS calls ahead
        self.host[i, :n] = torch.tensor(values, dtype=torch.int32)
        out = self.dev[i, :n]
        out.copy_(self.host[i, :n], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[i] = ev
        return out


_STAGERS = {}


def unknown_mask(alpha, widths, and_mask=None, alt=None, use_alt=None):
    """uint8 {0,1} mask of the dilated uncertain region (reference: utils/utils.py:28-55 compute_unknown).
    alpha [..., H, W] fp32; widths: one ellipse size (1..29) per [H, W] slice; and_mask optional uint8.
    alt / use_alt (optional): `use_alt` is a DEVICE int32 scalar; when it is non-zero the mask is computed from `alt`
    (same shape as alpha) instead - the reference's "guide with the ground truth when the predicted alpha is all zero"
    switch (decoder/resnet_inst_matt_spconv.py:311-316) without a host read."""
    _need_cuda(alpha, and_mask, alt, use_alt)
    a = alpha.detach().to(torch.float32).contiguous()
    H, W = a.shape[-2:]
    slices = a.numel() // (H * W) if a.numel() else 0
    out = torch.empty(a.shape, dtype=torch.uint8, device=a.device)
    if slices == 0:
        return out
    if len(widths) != slices:
        raise ValueError(f"unknown_mask: {len(widths)} widths for {slices} slices")
    w = _widths_tensor(tuple(int(v) for v in widths), a.device)
    if and_mask is not None:
        and_mask = and_mask.to(torch.uint8).contiguous()
        assert and_mask.shape == a.shape
    if alt is not None:
        b = alt.detach().to(torch.float32).contiguous()
        assert b.shape == a.shape and use_alt.dtype == torch.int32
        _lib.check(_lib.lib().mg_unknown_mask_select(_ptr(a), _ptr(b), _ptr(use_alt), slices, H, W, _ptr(w), _ptr(and_mask),
                                                    _ptr(out), None, _stream()), "mg_unknown_mask_select")
        return out
    _lib.check(_lib.lib().mg_unknown_mask(_ptr(a), slices, H, W, _ptr(w), _ptr(and_mask), _ptr(out), None, _stream()),
               "mg_unknown_mask")
    return out


def fuse_stage(src, finer, coarser, widths, and_mask):
    """One stage of the progressive fusion (reference: decoder/resnet_inst_matt_spconv.py:272-290 fuse()):
    w = unknown_mask(src, widths) & and_mask;  alpha = where(w, finer, coarser).  Returns (alpha fp32, w uint8).
    NATIVE (K10: the blend runs in the mask kernel's epilogue); no gradient flows through the fused alpha (the losses
    take the three scales separately), as in the reference's detached use."""
    _need_cuda(src, finer, coarser, and_mask)
    f = lambda t: t.detach().to(torch.float32).contiguous()
    s_, x, y = f(src), f(finer), f(coarser)
    H, W = s_.shape[-2:]
    slices = s_.numel() // (H * W) if s_.numel() else 0
    w_out = torch.empty(s_.shape, dtype=torch.uint8, device=s_.device)
    a_out = torch.empty_like(s_)
    if slices == 0:
        return a_out, w_out
    if len(widths) != slices:
        raise ValueError(f"fuse_stage: {len(widths)} widths for {slices} slices")
    wd = _widths_tensor(tuple(int(v) for v in widths), s_.device)
    am = and_mask.to(torch.uint8).contiguous()
    assert am.shape == s_.shape and x.shape == s_.shape and y.shape == s_.shape
    _lib.check(_lib.lib().mg_fuse_stage(_ptr(s_), _ptr(x), _ptr(y), slices, H, W, _ptr(wd), _ptr(am), _ptr(w_out), _ptr(a_out),
                                       _stream()), "mg_fuse_stage")
    return a_out, w_out


# =============================================================================================== native: K8b
SiteTables = namedtuple("SiteTables", "counts coords nbr parent child shapes flags", defaults=(None,))

# The step's status word: int32[8] on the device = four site counts + device-side flags, read by the host in ONE 32-byte
# copy (the step's only host read): [4] the predicted OS8 alpha is all zero (`x_os8.sum() == 0`,
# decoder/resnet_inst_matt_spconv.py:314), [5] some sample has no mask at all (the reference raises "Mask is empty",
# module/mask_attention.py:95-98).
STATUS_ALL_ZERO, STATUS_EMPTY_MASK = 4, 5
_STATUS_INIT = {}


def new_status(device):
    t = _STATUS_INIT.get(str(device))
    if t is None:
        t = _STATUS_INIT[str(device)] = torch.tensor([0, 0, 0, 0, 1, 0, 0, 0], dtype=torch.int32, device=device)
    return t.clone()


def build_sites(roi, status=None):
    """Active-site lists and rulebook tables of the 4 sparse levels (reference:
    decoder/resnet_inst_matt_spconv.py:203-218: nonzero + spconv dummy_downscale index generation).
    roi uint8 [slots, H, W].  One small D2H read: the four counts (+ the flags of `status`, see `new_status`, which then
    come back as `.flags`)."""
    _need_cuda(roi, status)
    roi = roi.to(torch.uint8).contiguous()
    S, H, W = roi.shape
    L = _lib.lib()
    ws = torch.empty(L.mg_sites_workspace(S, H, W), dtype=torch.uint8, device=roi.device)
    counts_d = status if status is not None else torch.zeros(4, dtype=torch.int32, device=roi.device)
    _lib.check(L.mg_sites_count(_ptr(roi), S, H, W, _ptr(ws), _ptr(counts_d), _stream()), "mg_sites_count")
    host = [int(c) for c in counts_d.cpu()]
    counts, flags = host[:4], (host[4:] if status is not None else None)
    if flags is not None and flags[STATUS_EMPTY_MASK - 4]:
        raise ValueError("Mask is empty")      # module/mask_attention.py:95-98
    mk = lambda n, k: torch.empty((n, k), dtype=torch.int32, device=roi.device)
    coords = [mk(n, 3) for n in counts]
    nbr = [mk(counts[0], 9), None, mk(counts[2], 9), None]
    parent = [mk(counts[l], 9) for l in range(3)] + [None]
    child = [None] + [mk(counts[l], 9) for l in range(1, 4)]
    pa = lambda ts: _lib.ptr_array([t.data_ptr() if (t is not None and t.numel()) else 0 for t in ts])
    _lib.check(L.mg_sites_tables(_ptr(ws), S, H, W, _lib.i32_array(counts), pa(coords), pa(nbr), pa(parent), pa(child),
                                 _stream()), "mg_sites_tables")
    shapes = [(H >> l, W >> l) for l in range(4)]
    return SiteTables(counts, coords, nbr, parent, child, shapes, flags)


# =============================================================================================== native: K1
def _mask_embed_f32(image, masks, table, slot_ids, C):
    """K1 with an fp32 NHWC output (fp32-accurate evaluation mode, no backward)."""
    _need_cuda(image, masks, table, slot_ids)
    image = image.detach().to(torch.float32).contiguous()
    masks = masks.detach().to(torch.float32).contiguous()
    B, _, H, W = image.shape
    out = torch.empty((B, H, W, C), dtype=torch.float32, device=image.device)
    tab = table.detach().to(torch.float32).contiguous()
    _lib.check(_lib.lib().mg_mask_embed_fwd_f32(_ptr(image), _ptr(masks), _ptr(slot_ids), masks.shape[1], _ptr(tab), _ptr(out),
                                               B, H, W, C, _stream()), "mg_mask_embed_fwd_f32")
    return out


class _MaskEmbed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, masks, table, slot_ids, C):
        _need_cuda(image, masks, table, slot_ids)
        image = image.to(torch.float32).contiguous()
        masks = masks.to(torch.float32).contiguous()
        B, _, H, W = image.shape
        M = masks.shape[1]
        out = torch.empty((B, H, W, C), dtype=torch.float16, device=image.device)
        tab = table.detach().to(torch.float32).contiguous()
        _lib.check(_lib.lib().mg_mask_embed_fwd(_ptr(image), _ptr(masks), _ptr(slot_ids), M, _ptr(tab), _ptr(out), B, H, W, C,
                                               _stream()), "mg_mask_embed_fwd")
        ctx.save_for_backward(masks, slot_ids)
        ctx.meta = (C, table.shape, table.dtype)
        return out

    @staticmethod
    def backward(ctx, gout):
        masks, slot_ids = ctx.saved_tensors
        C, tshape, tdtype = ctx.meta
        B, H, W, _ = gout.shape
        g = gout.to(torch.float16).contiguous()
        gtab = torch.zeros(tshape, dtype=torch.float32, device=gout.device)
        _lib.check(_lib.lib().mg_mask_embed_bwd(_ptr(g), _ptr(masks), _ptr(slot_ids), masks.shape[1], _ptr(gtab), B, H, W, C,
                                               _stream()), "mg_mask_embed_bwd")
        return None, None, gtab.to(tdtype), None, None


def slot_ids_tensor(slot_ids, device):
    """Device int32 tensor of the slot of each given mask (a tensor is passed through)."""
    if torch.is_tensor(slot_ids):
        return slot_ids.to(device=device, dtype=torch.int32)
    return _widths_tensor(tuple(int(s) for s in slot_ids), device)


def mask_embed(image, masks, table, slot_ids, C=8, dtype=torch.float16):
    """image [B,3,H,W] fp32, masks [B,M,H,W] {0,1}, table [11,3], slot_ids (list or device int32 tensor [M]) ->
    packed encoder input, NCHW-shaped channels-last fp16 [B,C,H,W] (ch 0-2 image, 3-5 mean id embedding, rest 0).
    dtype torch.float32: the fp32-accurate evaluation mode (no backward).
    Reference: arch/maggie.py:200-235 + encoder/resnet.py:211-229."""
    if tuple(table.shape) != (11, 3) or masks.shape[1] > 16:
        # K1 keeps the 11 x 3 id table of the reference configs (num_mask = 10, num_embed = 3) in shared memory
        raise NotImplementedError(f"mask_embed: id table must be [11, 3] and at most 16 masks per frame "
                                  f"(got table {tuple(table.shape)}, {masks.shape[1]} masks)")
    ids = slot_ids_tensor(slot_ids, image.device)
    if dtype == torch.float32:
        return _mask_embed_f32(image, masks, table, ids, C).permute(0, 3, 1, 2)
    return _MaskEmbed.apply(image, masks, table, ids, C).permute(0, 3, 1, 2)


# =============================================================================================== native: K12
class _MatteLossSums(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a1, a4, a8, target, w1, w4, w8, plane_scale):
        _need_cuda(a1, a4, a8, target, w1, w4, w8)
        f = lambda t: t.detach().to(torch.float32).contiguous()
        a1d, a4d, a8d, td, w1d, w4d, w8d = (f(t) for t in (a1, a4, a8, target, w1, w4, w8))
        psd = f(plane_scale).reshape(-1) if plane_scale is not None else None
        H, W = a1d.shape[-2:]
        S = a1d.numel() // (H * W)
        assert psd is None or psd.numel() == S
        L = _lib.lib()
        ws = torch.empty(L.mg_loss_workspace_floats(S, H, W), dtype=torch.float32, device=a1.device)
        n0 = 3 * S * H * W
        sg = torch.empty(n0 + n0 // 4 + n0 // 16, dtype=torch.float16, device=a1.device)
        from . import dense
        sums = dense.zeros_f32(32 * 3 * 8, a1.device).view(32, 3, 8)
        _lib.check(L.mg_loss_fwd(_ptr(a1d), _ptr(a4d), _ptr(a8d), _ptr(td), _ptr(w1d), _ptr(w4d), _ptr(w8d), _ptr(psd), S, H, W,
                                 _ptr(ws), _ptr(sg), _ptr(sums), _stream()), "mg_loss_fwd")
        ctx.save_for_backward(a1d, a4d, a8d, td, w1d, w4d, w8d, ws, sg, psd)
        ctx.shape = (S, H, W, a1.shape)
        return sums.sum(0)

    @staticmethod
    def backward(ctx, gs):
        a1d, a4d, a8d, td, w1d, w4d, w8d, ws, sg, psd = ctx.saved_tensors
        S, H, W, shape = ctx.shape
        coef = gs[:, :5].to(torch.float32).contiguous()
        g = torch.empty((3,) + tuple(a1d.shape), dtype=torch.float32, device=a1d.device)
        _lib.check(_lib.lib().mg_loss_bwd(_ptr(a1d), _ptr(a4d), _ptr(a8d), _ptr(td), _ptr(w1d), _ptr(w4d), _ptr(w8d), _ptr(psd), S,
                                         H, W, _ptr(ws), _ptr(sg), _ptr(coef), _ptr(g[0]), _ptr(g[1]), _ptr(g[2]), _stream()),
                   "mg_loss_bwd")
        return g[0].view(shape), g[1].view(shape), g[2].view(shape), None, None, None, None, None


def matte_loss_sums(a1, a4, a8, target, w1, w4, w8, plane_scale=None):
    """Partial sums [3 scales, 8] of the matting losses (weighted L1, Laplacian pyramid levels, Sobel gradient, weight
    sums) with a native backward to the three predictions.  plane_scale [planes] (optional): the predictions enter as
    a * plane_scale (the reference's `pred * valid_masks`).  NATIVE (K12).  Reference: arch/maggie.py:268-346."""
    return _MatteLossSums.apply(a1, a4, a8, target, w1, w4, w8, plane_scale)


# =============================================================================================== native: K0
def prepare_weights(bank):
    """Grouped weight preparation of every banked dense conv (spectral-norm power iteration, W / sigma, fp16 operand
    packs) with a grouped backward.  NATIVE (K0, maggie_b200/weights.py).  Reference: module/spectral_norm.py:22-35."""
    return bank.prepare()


def step_scope(key, device):
    """Per-forward scope (one zeroed scratch pool + deferred BatchNorm counters), see maggie_b200/dense.py."""
    from . import dense

    return dense.step_scope(key, device)


# =============================================================================================== interim ops
def _act(x, act):
    if act == "relu":
        return F.relu(x)
    if act == "lrelu":
        return F.leaky_relu(x, 0.2)
    assert act is None
    return x


def spectral_weight(w_bar, u, v):
    """One power iteration (updates u, v in place, no grad) and W = W_bar / sigma
    (reference: module/spectral_norm.py:22-35).  INTERIM (torch)."""
    h = w_bar.shape[0]
    with torch.no_grad():
        wm = w_bar.detach().reshape(h, -1)
        vn = torch.mv(wm.t(), u)
        vn = vn / (vn.norm() + 1e-12)
        un = torch.mv(wm, vn)
        un = un / (un.norm() + 1e-12)
        v.copy_(vn)
        u.copy_(un)
    # un / vn (fresh tensors) rather than the parameters: a layer may run several times per forward (video diff head)
    sigma = torch.dot(un, torch.mv(w_bar.reshape(h, -1), vn))
    return w_bar / sigma


def batch_norm(x, bn, training):
    """Training-mode (batch statistics, running-stat update) or eval-mode BatchNorm from a container's
    tensors.  INTERIM (torch)."""
    if training and bn.num_batches_tracked is not None:
        from . import dense
        dense.bump_counter(bn.num_batches_tracked)
    return F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, training, bn.momentum, bn.eps)


def conv_bn_act(x, w, bn, training, *, stride=1, padding=1, dilation=1, act="relu", act_first=False,
                residual=None, transposed=False, res_up=False):
    """conv (or 4x4 s2 transposed conv) -> BN -> (+residual) -> act, or conv -> act -> BN when act_first.
    NATIVE: tcgen05/TMA implicit-GEMM conv (K2) with BN-statistics epilogue, K3 BatchNorm kernels, native dgrad
    (K2 with the transposed pack) and wgrad (K4).  See maggie_b200/dense.py."""
    from . import dense

    return dense.conv_bn_act(x, w, bn, training, stride=stride, padding=padding, dilation=dilation, act=act,
                             act_first=act_first, residual=residual, transposed=transposed, res_up=res_up)


def conv_bias(x, w, bias=None, *, padding=1):
    """Plain conv + bias (ConvGRU gates, temporal-difference head).  Output channels are padded to the tensor-core
    granularity (16) here and sliced back.  NATIVE (K2 forward / dgrad, K4 wgrad)."""
    from . import dense

    co = w.shape[0]
    if co % 16:
        pad = 16 - co % 16
        w = F.pad(w, (0, 0, 0, 0, 0, 0, 0, pad))
        bias = F.pad(bias, (0, pad)) if bias is not None else None
    y = dense.conv_bias(x, w, bias, padding=padding)
    return y[:, :co] if y.shape[1] != co else y


def gru_step(x, h, w_ih, b_ih, w_hh, b_hh):
    """One ConvGRU step (module/conv_gru.py:50-58): rz = sigmoid(conv_ih([x | h])); c = tanh(conv_hh([x | r * h]));
    h' = (1 - z) h + z c.  ROUTED: fp16 CUDA tensors -> K11 gate kernels around the native convs (one autograd node, no
    torch.cat / sigmoid / tanh launches); anything else -> the torch composition."""
    C = x.shape[1]
    if x.is_cuda and x.dtype == torch.float16 and C % 8 == 0:
        from . import dense
        return dense.gru_step(x, h, w_ih, b_ih, w_hh, b_hh)
    rz = torch.sigmoid(conv_bias(torch.cat([x, h], 1), w_ih, b_ih).float())
    r, z = rz[:, :C], rz[:, C:]
    hf = h.float()
    c = torch.tanh(conv_bias(torch.cat([x, (r * hf).to(x.dtype)], 1), w_hh, b_hh).float())
    return ((1 - z) * hf + z * c).to(x.dtype)


class _TemporalFuse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fd, bd, preds):
        B, n_f, n_i, H, W = preds.shape
        fd, bd, preds = fd.contiguous(), bd.contiguous(), preds.contiguous()
        fused = torch.empty_like(preds)
        _lib.check(_lib.lib().mg_temporal_fuse_fwd(_ptr(fd), _ptr(bd), _ptr(preds), _ptr(fused), B, n_f, n_i, H * W, _stream()),
                   "mg_temporal_fuse_fwd")
        ctx.save_for_backward(fd, bd, preds)
        return fused

    @staticmethod
    def backward(ctx, g):
        fd, bd, preds = ctx.saved_tensors
        B, n_f, n_i, H, W = preds.shape
        g = g.contiguous().float()
        dp, dfd, dbd = torch.empty_like(preds), torch.empty_like(fd), torch.empty_like(bd)
        _lib.check(_lib.lib().mg_temporal_fuse_bwd(_ptr(fd), _ptr(bd), _ptr(preds), _ptr(g), _ptr(dp), _ptr(dfd), _ptr(dbd), B, n_f,
                                                   n_i, H * W, _stream()), "mg_temporal_fuse_bwd")
        return dfd, dbd, dp


def temporal_fuse(fd, bd, preds):
    """Bidirectional alpha fusion over the frames of a clip (decoder/resnet_inst_matt_spconv_temp.py:122-142).
    fd, bd [B, n_f, 1, H, W] fp32 logits of the forward / backward temporal-difference maps (fd[:, 0] and bd[:, -1] are the
    reference's zero planes and are not read); preds [B, n_f, n_i, H, W] fp32.  ROUTED: CUDA fp32 with n_f <= 8 -> K11 (one
    pass forward, one backward); anything else -> the torch recurrences."""
    n_f = preds.shape[1]
    if preds.is_cuda and preds.dtype == fd.dtype == bd.dtype == torch.float32 and 2 <= n_f <= 8:
        return _TemporalFuse.apply(fd, bd, preds)
    sf, sb = torch.sigmoid(fd), torch.sigmoid(bd)
    fp = [preds[:, 0]]
    for i in range(1, n_f):
        fp.append(fp[-1] * (1 - sf[:, i]) + preds[:, i] * sf[:, i])
    bp = [preds[:, n_f - 1]]
    for i in range(n_f - 2, -1, -1):
        bp.append(bp[-1] * (1 - sb[:, i]) + preds[:, i] * sb[:, i])
    bp = bp[::-1]
    return torch.stack([fp[0]] + [(fp[i] + bp[i]) / 2 for i in range(1, n_f - 1)] + [bp[n_f - 1]], 1)


def linear(x, w, b=None):
    return F.linear(x, w.to(x.dtype), None if b is None else b.to(x.dtype))


class _SplitRows(torch.autograd.Function):
    """w [n * E, ...] -> n row chunks (views).  The backward is ONE concatenation instead of the zeros + copy + add chain
    autograd builds for `w[:E]`, `w[E:2E]`, ... (the q / k / v slices of an in_proj matrix: ~8 launches per matrix)."""

    @staticmethod
    def forward(ctx, w, n):
        E = w.shape[0] // n
        ctx.meta = (n, E, tuple(w.shape[1:]), w.dtype, w.device)
        return tuple(w.narrow(0, i * E, E) for i in range(n))

    @staticmethod
    def backward(ctx, *gs):
        n, E, rest, dt, dev = ctx.meta
        if all(g is None for g in gs):
            return None, None
        parts = [g.to(dt) if g is not None else torch.zeros((E,) + rest, dtype=dt, device=dev) for g in gs]
        return torch.cat(parts, 0), None


def split_rows(w, n):
    """The n equal row chunks of w as views (q / k / v parts of `in_proj_weight` / `in_proj_bias`)."""
    return _SplitRows.apply(w, n)


def small_linear(x, w, b=None, pos=None, relu=False):
    """Linear layer on the TOKEN side (10 instance tokens per sample, a few hundred fp32 rows of <= 128 features):
    y = act((x + pos) W^T + b).  A plain library GEMM (cuBLAS through torch) on the fp32 master weights - no casts, no
    weight copies.  (A hand-written fp32 kernel with shared-memory-staged operands, one launch per direction, was measured
    at 13 us against 7 us for add + addmm on the [80 x 128] x [128 x 128] layer and cost 0.45 ms per C2 step in an A/B
    run: these 1-MFLOP GEMMs are pure launch latency, and the library's small-GEMM kernels win.)"""
    if pos is not None:
        x = x + pos
    y = linear(x, w, b)
    return F.relu(y) if relu else y


def linear_rows(x, w, b=None, pos=None):
    """Linear layer on [..., Cin] rows; `pos` (optional) is added to the input first (positional terms of Q / K).
    Large row counts (the pixel side of the attention: B*4096 rows) run on the tcgen05 rows GEMM with native gradients
    (K9 with T = 1); the 10-token side is a library GEMM on the fp32 master weights (`small_linear`)."""
    rows = x.numel() // x.shape[-1]
    if rows < 1024 or not x.is_cuda:
        return small_linear(x, w, b, pos=pos)
    if pos is not None:
        x = x + pos
    if x.dtype == torch.float32:      # fp32-accurate evaluation mode: split-operand tensor-core GEMM
        from . import dense
        return dense.linear_rows_x3(x, w, b)
    y = rows_conv(x.reshape(rows, x.shape[-1]), w, b)
    return y.reshape(*x.shape[:-1], w.shape[0]).to(x.dtype)


class _LayerNormRes(torch.autograd.Function):
    """LayerNorm(a + b) on fp16 rows of 64 / 128 elements, fp32 statistics (K13)."""

    @staticmethod
    def forward(ctx, a, b, gamma, beta, eps):
        E = a.shape[-1]
        rows = a.numel() // E
        ah = a.detach().contiguous()
        bh = b.detach().contiguous() if b is not None else None
        y = torch.empty_like(ah)
        s = torch.empty_like(ah) if bh is not None else ah
        stat = torch.empty((rows, 2), dtype=torch.float32, device=a.device)
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        _lib.check(_lib.lib().mg_layer_norm_fwd(_ptr(ah), _ptr(bh), _ptr(g32), _ptr(b32), float(eps),
                                               _ptr(s) if bh is not None else None, _ptr(y), _ptr(stat), rows, E, _stream()),
                   "mg_layer_norm_fwd")
        ctx.save_for_backward(s, g32, stat)
        ctx.meta = (rows, E, b is not None, gamma.dtype)
        return y

    @staticmethod
    def backward(ctx, gy):
        s, g32, stat = ctx.saved_tensors
        rows, E, has_b, pdt = ctx.meta
        g = gy.contiguous() if gy.dtype == torch.float16 else gy.to(torch.float16).contiguous()
        dx = torch.empty_like(s)
        from . import dense
        dgb = dense.zeros_f32(2 * E, s.device).view(2, E)
        _lib.check(_lib.lib().mg_layer_norm_bwd(_ptr(s), _ptr(g), _ptr(g32), _ptr(stat), _ptr(dx), _ptr(dgb), rows, E, _stream()),
                   "mg_layer_norm_bwd")
        return dx, (dx if has_b else None), dgb[0].to(pdt), dgb[1].to(pdt), None


def layer_norm(x, ln, residual=None):
    """LayerNorm(x + residual) from a container's tensors.  NATIVE (K13, residual add fused) for CUDA fp16 rows of 64 / 128
    elements; torch composition otherwise (fp32 inputs, CPU tests)."""
    if x.is_cuda and x.dtype == torch.float16 and x.shape[-1] in (64, 128) and (residual is None or residual.dtype == x.dtype):
        if residual is not None and residual.shape != x.shape:
            residual = residual.expand_as(x)
        return _LayerNormRes.apply(x, residual, ln.weight, ln.bias, ln.eps)
    if (x.is_cuda and x.dtype == torch.float32 and x.shape[-1] in (64, 128) and x.numel() // x.shape[-1] >= 1024
            and not torch.is_grad_enabled()):
        # fp32-accurate evaluation mode, pixel rows (the 10-token side stays a tiny torch op)
        a = x.contiguous()
        r = residual.to(torch.float32).expand_as(x).contiguous() if residual is not None else None
        y = torch.empty_like(a)
        E = x.shape[-1]
        _lib.check(_lib.lib().mg_layer_norm_fwd_f32(_ptr(a), _ptr(r), _ptr(ln.weight.detach().float().contiguous()),
                                                   _ptr(ln.bias.detach().float().contiguous()), float(ln.eps), _ptr(y),
                                                   a.numel() // E, E, _stream()), "mg_layer_norm_fwd_f32")
        return y
    if residual is not None:
        x = x + residual
    return F.layer_norm(x.float(), (x.shape[-1],), ln.weight, ln.bias, ln.eps).to(x.dtype)


class _TokenLogits(torch.autograd.Function):
    """logits[bt, q, h, w] = sum_c tok[bt // n_f, q, c] * x[bt, c, h, w] with x channels-last fp16 (K13)."""

    @staticmethod
    def forward(ctx, tok, x, n_f):
        xn = x.permute(0, 2, 3, 1)
        assert xn.is_contiguous() and xn.dtype == torch.float16
        BT, H, W, C = xn.shape
        tk = tok.detach().to(torch.float32).contiguous()
        Q = tk.shape[1]
        out = torch.empty((BT, Q, H, W), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().mg_token_logits_fwd(_ptr(tk), _ptr(xn), _ptr(out), BT, n_f, Q, H * W, C, _stream()),
                   "mg_token_logits_fwd")
        ctx.save_for_backward(tk, xn)
        ctx.meta = (n_f, tok.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        tk, xn = ctx.saved_tensors
        n_f, tdt = ctx.meta
        BT, H, W, C = xn.shape
        Q = tk.shape[1]
        gc = g.to(torch.float32).contiguous()
        dx = torch.empty_like(xn) if ctx.needs_input_grad[1] else None
        from . import dense
        dtok = dense.zeros_f32(tk.numel(), tk.device).view(tk.shape) if ctx.needs_input_grad[0] else None
        _lib.check(_lib.lib().mg_token_logits_bwd(_ptr(tk), _ptr(xn), _ptr(gc), _ptr(dx), _ptr(dtok), BT, n_f, Q, H * W, C,
                                                 _stream()), "mg_token_logits_bwd")
        return (dtok.to(tdt) if dtok is not None else None), (dx.permute(0, 3, 1, 2) if dx is not None else None), None


def token_logits(tok, x, n_f):
    """The OS8 head's `einsum('bqc,btchw->btqhw')` (instance_matte_decoder.py:302) flattened to [b*n_f, q, h, w], fp32.
    tok [b, q, 64] fp32; x [b*n_f, 64, h, w].  NATIVE (K13) on CUDA channels-last fp16 features; torch otherwise."""
    if x.is_cuda and x.dtype == torch.float16 and x.shape[1] == 64 and tok.shape[1] <= 16 and x.permute(0, 2, 3, 1).is_contiguous():
        return _TokenLogits.apply(tok, x, n_f)
    if (x.is_cuda and x.dtype == torch.float32 and x.shape[1] == 64 and tok.shape[1] <= 16 and not torch.is_grad_enabled()):
        xn = x.permute(0, 2, 3, 1).contiguous()      # fp32-accurate evaluation mode
        BT, H, W, C = xn.shape
        tk = tok.detach().to(torch.float32).contiguous()
        out = torch.empty((BT, tk.shape[1], H, W), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().mg_token_logits_fwd_f32(_ptr(tk), _ptr(xn), _ptr(out), BT, n_f, tk.shape[1], H * W, C, _stream()),
                   "mg_token_logits_fwd_f32")
        return out
    b = tok.shape[0]
    return torch.einsum("bqc,btchw->btqhw", tok, x.float().reshape(b, n_f, *x.shape[1:])).flatten(0, 1)


class _IdEmbedding(torch.autograd.Function):
    """table[ids] for a tiny table (11 id rows) and many ids (b * 4096 * n_f pixels).  The forward is a plain gather; the
    backward is ONE one-hot GEMM (table^T-shaped: [rows, ids] x [ids, E], fp32) instead of torch's sort-based
    embedding_dense_backward (128 us at C2 for a [11, 128] gradient)."""

    @staticmethod
    def forward(ctx, ids, table, dtype):
        flat = ids.reshape(-1)
        ctx.save_for_backward(flat)
        ctx.meta = (table.shape[0], table.dtype)
        return table.detach().index_select(0, flat).view(*ids.shape, table.shape[1]).to(dtype)

    @staticmethod
    def backward(ctx, g):
        (flat,) = ctx.saved_tensors
        rows, tdt = ctx.meta
        onehot = F.one_hot(flat, (rows + 7) // 8 * 8).to(torch.float32)             # [ids, rows padded to 8]
        grad = onehot.t().matmul(g.reshape(flat.numel(), -1).to(torch.float32))     # [rows_pad, E]
        return None, grad[:rows].to(tdt), None


def id_embedding(ids, table, dtype):
    """F.embedding(ids, table).to(dtype) with a GEMM backward (see _IdEmbedding); ids int64 [...], table [rows, E]."""
    return _IdEmbedding.apply(ids, table, dtype)


def col_sum(x):
    """fp32 column sums of fp16 rows [N, C] (bias gradients of the sparse layers).  NATIVE (K13)."""
    _need_cuda(x)
    N, C = x.shape
    if x.dtype != torch.float16 or x.stride(1) != 1 or x.stride(0) % 8 or C % 8 or 256 % (C // 8):
        return x.float().sum(0)
    from . import dense
    out = dense.zeros_f32(C, x.device)
    _lib.check(_lib.lib().mg_col_sum(_ptr(x), x.stride(0), N, C, _ptr(out), _stream()), "mg_col_sum")
    return out


class _AttnTQ(torch.autograd.Function):
    """Few queries (tokens) over many keys, with the attention-max statistic (K6 "tq")."""

    @staticmethod
    def forward(ctx, q, k, v, key_pad, guidance):
        _need_cuda(q, k, v)
        B, Fq, E = q.shape
        S = k.shape[1]
        qf = q.detach().to(torch.float32).contiguous()
        kh, vh = k.detach().to(torch.float16).contiguous(), v.detach().to(torch.float16).contiguous()
        kp = key_pad.to(torch.uint8).contiguous() if key_pad is not None else None
        gd = guidance.to(torch.uint8).contiguous() if guidance is not None else None
        L = _lib.lib()
        dev = q.device
        out = torch.empty((B, Fq, E), dtype=torch.float32, device=dev)
        small = torch.empty((3, B, Fq), dtype=torch.float32, device=dev)   # stat, row max, row sum
        ws = torch.empty(L.mg_attn_tq_workspace_floats(B, Fq, S), dtype=torch.float32, device=dev)
        _lib.check(L.mg_attn_tq_fwd(_ptr(qf), _ptr(kh), _ptr(vh), _ptr(kp), _ptr(gd), B, Fq, S, E, _ptr(out), _ptr(small[0]),
                                    _ptr(small[1]), _ptr(small[2]), _ptr(ws), _stream()), "mg_attn_tq_fwd")
        ctx.save_for_backward(qf, kh, vh, kp, gd, out, small)
        ctx.dt = (q.dtype, k.dtype, v.dtype)
        return out.to(q.dtype), small[0]

    @staticmethod
    def backward(ctx, g_out, g_stat):
        qf, kh, vh, kp, gd, out, small = ctx.saved_tensors
        B, Fq, E = qf.shape
        S = kh.shape[1]
        go = g_out.to(torch.float32).contiguous()
        gs = g_stat.to(torch.float32).contiguous() if (g_stat is not None and gd is not None) else None
        from . import dense
        dq = dense.zeros_f32(qf.numel(), qf.device).view(qf.shape)
        dk, dv = torch.empty_like(kh), torch.empty_like(vh)
        _lib.check(_lib.lib().mg_attn_tq_bwd(_ptr(qf), _ptr(kh), _ptr(vh), _ptr(kp), _ptr(gd), _ptr(out), _ptr(small[0]),
                                            _ptr(small[1]), _ptr(small[2]), _ptr(go), _ptr(gs), B, Fq, S, E, _ptr(dq),
                                            _ptr(dk), _ptr(dv), _stream()), "mg_attn_tq_bwd")
        qd, kd, vd = ctx.dt
        return dq.to(qd), dk.to(kd), dv.to(vd), None, None


class _AttnFQ(torch.autograd.Function):
    """Many queries (pixels) over few keys (tokens, with key padding) (K6 "fq")."""

    @staticmethod
    def forward(ctx, q, k, v, key_pad):
        _need_cuda(q, k, v)
        B, S, E = q.shape
        Fk = k.shape[1]
        qh = q.detach().to(torch.float16).contiguous()
        kf, vf = k.detach().to(torch.float32).contiguous(), v.detach().to(torch.float32).contiguous()
        kp = key_pad.to(torch.uint8).contiguous() if key_pad is not None else None
        out = torch.empty((B, S, E), dtype=torch.float16, device=q.device)
        _lib.check(_lib.lib().mg_attn_fq_fwd(_ptr(qh), _ptr(kf), _ptr(vf), _ptr(kp), B, Fk, S, E, _ptr(out), _stream()),
                   "mg_attn_fq_fwd")
        ctx.save_for_backward(qh, kf, vf, kp)
        ctx.dt = (q.dtype, k.dtype, v.dtype)
        return out.to(q.dtype)

    @staticmethod
    def backward(ctx, g_out):
        qh, kf, vf, kp = ctx.saved_tensors
        B, S, E = qh.shape
        Fk = kf.shape[1]
        go = g_out.to(torch.float16).contiguous()
        dq = torch.empty_like(qh)
        from . import dense
        dk = dense.zeros_f32(kf.numel(), kf.device).view(kf.shape)
        dv = dense.zeros_f32(vf.numel(), vf.device).view(vf.shape)
        _lib.check(_lib.lib().mg_attn_fq_bwd(_ptr(qh), _ptr(kf), _ptr(vf), _ptr(kp), _ptr(go), B, Fk, S, E, _ptr(dq), _ptr(dk),
                                            _ptr(dv), _stream()), "mg_attn_fq_bwd")
        qd, kd, vd = ctx.dt
        return dq.to(qd), dk.to(kd), dv.to(vd), None


def attention(q, k, v, key_padding=None, need_stat=None):
    """Single-head attention, batch-first: q [B,L,E], k/v [B,S,E] (already projected), softmax in fp32.
    key_padding [B,S] bool (True = ignore).  need_stat: optional [B,L,S] bool guidance mask; if given also returns
    stat[b,l] = sum_s guidance * A (the only thing the attention-max loss needs, instance_matte_decoder.py:101-109).
    NATIVE (K6): few-query ("tq") kernel when L <= 16, many-query ("fq") kernel when S <= 16."""
    hp = not torch.is_grad_enabled() and (k.dtype == torch.float32 if q.shape[1] <= 16 else q.dtype == torch.float32)
    if hp:
        return _attention_f32(q, k, v, key_padding, need_stat)
    if q.shape[1] <= 16:
        o, stat = _AttnTQ.apply(q, k, v, key_padding, need_stat)
        return o, (stat if need_stat is not None else None)
    assert k.shape[1] <= 16 and need_stat is None, "attention: one side must have <= 16 rows"
    return _AttnFQ.apply(q, k, v, key_padding), None


def _attention_f32(q, k, v, key_padding=None, need_stat=None):
    """The attention cores with every operand in fp32 (fp32-accurate evaluation mode, forward only)."""
    _need_cuda(q, k, v)
    f = lambda t: t.detach().to(torch.float32).contiguous()
    u8 = lambda t: t.to(torch.uint8).contiguous() if t is not None else None
    qf, kf, vf, kp = f(q), f(k), f(v), u8(key_padding)
    L, dev = _lib.lib(), q.device
    if q.shape[1] <= 16:
        B, Fq, E = qf.shape
        S = kf.shape[1]
        gd = u8(need_stat)
        out = torch.empty((B, Fq, E), dtype=torch.float32, device=dev)
        small = torch.empty((3, B, Fq), dtype=torch.float32, device=dev)
        ws = torch.empty(L.mg_attn_tq_workspace_floats(B, Fq, S), dtype=torch.float32, device=dev)
        _lib.check(L.mg_attn_tq_fwd_f32(_ptr(qf), _ptr(kf), _ptr(vf), _ptr(kp), _ptr(gd), B, Fq, S, E, _ptr(out), _ptr(small[0]),
                                        _ptr(small[1]), _ptr(small[2]), _ptr(ws), _stream()), "mg_attn_tq_fwd_f32")
        return out, (small[0] if need_stat is not None else None)
    assert k.shape[1] <= 16 and need_stat is None, "attention: one side must have <= 16 rows"
    B, S, E = qf.shape
    out = torch.empty((B, S, E), dtype=torch.float32, device=dev)
    _lib.check(L.mg_attn_fq_fwd_f32(_ptr(qf), _ptr(kf), _ptr(vf), _ptr(kp), B, kf.shape[1], S, E, _ptr(out), _stream()),
               "mg_attn_fq_fwd_f32")
    return out, None


def rows_conv(src, w, bias=None, *, table=None, table_t=None, mirror=False, bn=None, mode="plain", act=None,
              training=False):
    """Rulebook convolution on site rows (+ BatchNorm1d + activation), NATIVE (K9 gather -> tcgen05 tile -> rows;
    K3 BatchNorm kernels; native data / weight gradients).  See maggie_b200/sparse.py."""
    from . import sparse

    return sparse.rows_conv(src, w, bias, table=table, table_t=table_t, mirror=mirror, bn=bn, mode=mode, act=act,
                            training=training)


def rows_head(src, w, bias, nbr, coords, slots, H, W):
    """SubM 3x3 C->1 head written into the fp32 logit map (-99 where inactive), NATIVE (K9)."""
    from . import sparse

    return sparse.rows_head(src, w, bias, nbr, coords, slots, H, W)


def gather_dense(dense, coords, n_i):
    """dense NCHW-shaped channels-last [B,C,H,W] -> rows [N,C] at coords (frame = slot // n_i), NATIVE (K9)."""
    from . import sparse

    return sparse.gather_dense(dense, coords, n_i)


class _UpsampleTanh(torch.autograd.Function):
    """(tanh(bilinear_upsample(logits)) + 1) / 2 * plane_scale, fp32 (K7)."""

    @staticmethod
    def forward(ctx, logits, plane_scale, S, all_zero=None):
        x = logits.detach().to(torch.float32).contiguous()
        h, w = x.shape[-2:]
        planes = x.numel() // (h * w) if x.numel() else 0
        ps = plane_scale.detach().to(torch.float32).contiguous() if plane_scale is not None else None
        out = torch.empty(x.shape[:-2] + (h * S, w * S), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().mg_upsample_tanh_fwd(_ptr(x), _ptr(ps), _ptr(out), planes, h, w, S, _ptr(all_zero), _stream()),
                   "mg_upsample_tanh_fwd")
        ctx.save_for_backward(x, ps)
        ctx.meta = (planes, h, w, S, logits.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        x, ps = ctx.saved_tensors
        planes, h, w, S, dt = ctx.meta
        gl = torch.empty_like(x)
        _lib.check(_lib.lib().mg_upsample_tanh_bwd(_ptr(x), _ptr(ps), _ptr(g.to(torch.float32).contiguous()), _ptr(gl), planes, h, w,
                                                  S, _stream()), "mg_upsample_tanh_bwd")
        return gl.to(dt), None, None, None


def upsample_tanh(logits, size=None, scale=None, plane_scale=None, all_zero=None):
    """bilinear (align_corners=False) -> (tanh+1)/2 (-> * plane_scale [planes]) in fp32.  NATIVE (K7) on CUDA for the integer
    scales 1 / 2 / 4 / 8 the path uses; torch composition otherwise (host goldens).
    all_zero (optional): device int32 scalar preset to 1, cleared by the kernel when some output is non-zero."""
    S = 1
    if size is not None:
        S = size[-1] // logits.shape[-1] if size[-1] % logits.shape[-1] == 0 and size[-2] == logits.shape[-2] * (size[-1] // logits.shape[-1]) else 0
    elif scale is not None:
        S = int(scale) if float(scale) == int(scale) else 0
    if logits.is_cuda and S in (1, 2, 4, 8) and logits.shape[-3:].numel() > 0:
        return _UpsampleTanh.apply(logits, plane_scale, S, all_zero)
    x = logits.float()
    if size is not None or scale is not None:
        x = F.interpolate(x, size=size, scale_factor=scale, mode="bilinear", align_corners=False)
    a = (torch.tanh(x) + 1.0) / 2.0
    if plane_scale is not None:
        a = a * plane_scale.reshape(a.shape[:-2] + (1, 1)).to(a.dtype)
    if all_zero is not None:
        all_zero.copy_((a.detach().amax() == 0).to(all_zero.dtype))
    return a
