"""Native dense ops (K2 conv on tcgen05/TMA, ...) - python side of the C ABI: descriptor structs, weight packs,
tap tables.  Tensors are NHWC fp16 (`[N,H,W,C]` contiguous)."""
import ctypes
import os
from ctypes import c_int32, c_void_p

import torch
import torch.nn.functional as F

from . import _lib
from .weights import BankedWeight
_need_cuda, _ptr, _stream = _lib.need_cuda, _lib.tensor_ptr, _lib.stream_ptr


def _banked(w):
    return isinstance(w, BankedWeight)


def wshape(w):
    """Torch-layout shape of a weight tensor, or the equivalent for a bank handle ([Ci,Co,4,4] when transposed)."""
    if _banked(w):
        Co, Ci, kh, kw = w.logical
        return (Ci, Co, kh, kw) if w.transposed else (Co, Ci, kh, kw)
    return tuple(w.shape)

MAX_TAPS = 16
STAT_COPIES = 16
FUSED_BN_APPLY = os.environ.get("MAGGIE_B200_NO_FUSED_BN_APPLY", "0") != "1"   # K3: finalize + apply in one launch (training)


class ConvDesc(ctypes.Structure):
    """Mirror of `mg_conv_desc` (include/maggie_b200.h)."""
    _fields_ = [
        ("x", c_void_p), ("N", c_int32), ("Hi", c_int32), ("Wi", c_int32), ("Ci", c_int32),
        ("w", c_void_p), ("Co", c_int32), ("Ktot", c_int32),
        ("n_taps", c_int32), ("tap_dy", c_int32 * MAX_TAPS), ("tap_dx", c_int32 * MAX_TAPS), ("tap_koff", c_int32 * MAX_TAPS),
        ("sy", c_int32), ("sx", c_int32), ("Hg", c_int32), ("Wg", c_int32),
        ("out", c_void_p), ("Ho", c_int32), ("Wo", c_int32), ("Cs", c_int32), ("c_off", c_int32),
        ("oys", c_int32), ("oy0", c_int32), ("oxs", c_int32), ("ox0", c_int32),
        ("pre_act", c_int32), ("post_act", c_int32), ("stats", c_void_p), ("bias", c_void_p),
        ("scale", c_void_p), ("shift", c_void_p), ("res", c_void_p), ("res_up", c_int32),
        ("n_phases", c_int32), ("phase_tap0", c_int32 * 5), ("phase_oy0", c_int32 * 4), ("phase_ox0", c_int32 * 4),
    ]

ACT = {None: 0, "relu": 1, "lrelu": 2}

def pad_channels(c):
    return (c + 15) // 16 * 16


def pack_weight(w, ci_pad=None):
    """[Co,Ci,kh,kw] -> fp16 [Co, kh*kw*Ci_pad] (tap-major, channel contiguous)."""
    Co, Ci, kh, kw = w.shape
    ci_pad = ci_pad or pad_channels(Ci)
    p = w.permute(0, 2, 3, 1)
    if ci_pad != Ci:
        p = F.pad(p, (0, ci_pad - Ci))
    return p.reshape(Co, kh * kw * ci_pad).to(torch.float16).contiguous()


def pack_weight_f32(w, ci_pad=None):
    """[Co,Ci,kh,kw] -> fp32 [Co, kh*kw*Ci_pad] (the layout of `pack_weight`, unrounded: input of `split_f32`)."""
    Co, Ci, kh, kw = w.shape
    ci_pad = ci_pad or pad_channels(Ci)
    p = w.detach().float().permute(0, 2, 3, 1)
    if ci_pad != Ci:
        p = F.pad(p, (0, ci_pad - Ci))
    return p.reshape(Co, kh * kw * ci_pad).contiguous()


def split_f32(t):
    """fp32 tensor -> (hi, lo) fp16 tensors with t = hi + lo up to 2^-22 |t| (K17): the operand form of the fp32-accurate
    ("x3") evaluation mode of the tensor-core kernels."""
    _need_cuda(t)
    t = t.contiguous()
    assert t.dtype == torch.float32 and t.numel() % 4 == 0
    hi, lo = torch.empty_like(t, dtype=torch.float16), torch.empty_like(t, dtype=torch.float16)
    _lib.check(_lib.lib().mg_split_f32(_ptr(t), _ptr(hi), _ptr(lo), t.numel(), _stream()), "mg_split_f32")
    return hi, lo


def conv_taps(kh, kw, pad, dil, ci_pad):
    """Forward-conv tap table: [(dy, dx, koff)]."""
    return [(ky * dil - pad, kx * dil - pad, (ky * kw + kx) * ci_pad) for ky in range(kh) for kx in range(kw)]


def conv_launch(x, w_packed, taps, *, stride=1, grid_hw=None, out=None, out_hw=None, out_map=(1, 0, 1, 0), c_off=0,
                relu=False, stats=None, bias=None, pre_act=None, post_act=None, scale=None, shift=None, res=None,
                res_up=False, phases=None, lo=None):
    """Generic launch of mg_conv_fprop.  x [N,Hi,Wi,Ci] fp16 NHWC; w_packed [Co,Ktot] fp16; taps [(dy,dx,koff)].
    grid_hw: logical output grid (defaults to ceil(Hi/stride)); out: preallocated NHWC fp16 (or None);
    out_map = (oys, oy0, oxs, ox0).
    phases: [(taps, oy0, ox0), ...] (2..4 entries, `taps` ignored): the sub-pixel phases of a stride-2 data gradient /
    transposed conv as ONE launch; out_map supplies the common output strides (oys, _, oxs, _).
    lo = (x_lo, w_lo): fp32-accurate "x3" mode (mg_conv_fprop_x3): x / w_packed are the hi halves of split operands, the
    output and the residual are FP32 tensors."""
    if phases is not None:
        taps = [t for ph in phases for t in ph[0]]
    _need_cuda(x, w_packed)
    assert x.dtype == torch.float16 and x.is_contiguous() and w_packed.dtype == torch.float16 and w_packed.is_contiguous()
    N, Hi, Wi, Ci = x.shape
    Co, Ktot = w_packed.shape
    Hg, Wg = grid_hw if grid_hw is not None else ((Hi + stride - 1) // stride, (Wi + stride - 1) // stride)
    odt = torch.float32 if lo is not None else torch.float16
    if out is None:
        Ho, Wo = out_hw if out_hw is not None else (Hg, Wg)
        out = torch.empty((N, Ho, Wo, Co), dtype=odt, device=x.device)
    assert out.dtype == odt and out.is_contiguous()
    assert res is None or (res.dtype == odt and res.is_contiguous())
    d = ConvDesc()
    d.x, d.N, d.Hi, d.Wi, d.Ci = x.data_ptr(), N, Hi, Wi, Ci
    d.w, d.Co, d.Ktot = w_packed.data_ptr(), Co, Ktot
    d.n_taps = len(taps)
    for i, (dy, dx, ko) in enumerate(taps):
        d.tap_dy[i], d.tap_dx[i], d.tap_koff[i] = dy, dx, ko
    d.sy = d.sx = stride
    d.Hg, d.Wg = Hg, Wg
    d.out, d.Ho, d.Wo, d.Cs, d.c_off = out.data_ptr(), out.shape[1], out.shape[2], out.shape[3], c_off
    d.oys, d.oy0, d.oxs, d.ox0 = out_map
    d.pre_act = ACT["relu" if relu else pre_act]
    d.post_act = ACT[post_act]
    d.stats = stats.data_ptr() if stats is not None else None
    d.bias = bias.data_ptr() if bias is not None else None
    d.scale = scale.data_ptr() if scale is not None else None
    d.shift = shift.data_ptr() if shift is not None else None
    d.res = res.data_ptr() if res is not None else None
    d.res_up = int(res_up)
    if phases is not None:
        d.n_phases = len(phases)
        t0 = 0
        for i, (ptaps, oy0, ox0) in enumerate(phases):
            d.phase_tap0[i], d.phase_oy0[i], d.phase_ox0[i] = t0, oy0, ox0
            t0 += len(ptaps)
        d.phase_tap0[len(phases)] = t0
    if lo is not None:
        x_lo, w_lo = lo
        assert x_lo.shape == x.shape and x_lo.dtype == torch.float16 and x_lo.is_contiguous()
        assert w_lo.shape == w_packed.shape and w_lo.dtype == torch.float16 and w_lo.is_contiguous() and stats is None
        _lib.check(_lib.lib().mg_conv_fprop_x3(ctypes.byref(d), _ptr(x_lo), _ptr(w_lo), _stream()), "mg_conv_fprop_x3")
        return out
    _lib.check(_lib.lib().mg_conv_fprop(ctypes.byref(d), _stream()), "mg_conv_fprop")
    return out


def conv2d_nhwc(x, w, *, stride=1, padding=1, dilation=1, relu=False, stats=None, bias=None, out=None, c_off=0, **epi):
    """Forward conv: x NHWC fp16 (channels already padded to a multiple of 16), w torch layout [Co,Ci,kh,kw]."""
    Co, Ci, kh, kw = wshape(w)
    ci_pad = x.shape[-1]
    if _banked(w):
        assert ci_pad == w.ci_pad
        wp = w.P
    else:
        wp = pack_weight(w, ci_pad)
    Hi, Wi = x.shape[1:3]
    Ho = (Hi + 2 * padding - dilation * (kh - 1) - 1) // stride + 1
    Wo = (Wi + 2 * padding - dilation * (kw - 1) - 1) // stride + 1
    return conv_launch(x, wp, conv_taps(kh, kw, padding, dilation, ci_pad), stride=stride, grid_hw=(Ho, Wo), relu=relu,
                       stats=stats, bias=bias, out=out, c_off=c_off, **epi)


def conv_transpose4x4s2_nhwc(x, w, *, stats=None):
    """ConvTranspose2d(k=4, s=2, p=1): w torch layout [Ci,Co,4,4]; four sub-pixel phase convs (2x2 taps each).
    out[2y+py, 2x+px] = sum_{ky,kx} w[:, :, ky, kx]^T x[y+dy, x+dx] with dy = (py + 1 - ky) / 2 (integer)."""
    Ci, Co, kh, kw = wshape(w)
    assert kh == 4 and kw == 4
    N, Hi, Wi, ci_pad = x.shape
    out = torch.empty((N, 2 * Hi, 2 * Wi, Co), dtype=torch.float16, device=x.device)
    wp = w.P if _banked(w) else pack_weight(w.permute(1, 0, 2, 3), ci_pad)  # [Co][ky][kx][Ci]
    conv_launch(x, wp, None, grid_hw=(Hi, Wi), out=out, out_map=(2, 0, 2, 0), stats=stats, phases=convT_phases(ci_pad))
    return out


def convT_phases(ci_pad):
    """The four sub-pixel phases of ConvTranspose2d(4, 2, 1): [(taps, oy0, ox0)], 2 x 2 taps each."""
    phases = []
    for py in range(2):
        for px in range(2):
            taps = [((py + 1 - ky) // 2, (px + 1 - kx) // 2, (ky * 4 + kx) * ci_pad)
                    for ky in range(4) if (py + 1 - ky) % 2 == 0 for kx in range(4) if (px + 1 - kx) % 2 == 0]
            phases.append((taps, py, px))
    return phases


# ---- per-forward scope: one zeroed scratch pool + deferred BatchNorm counters -----------------------------------
# Every training conv needs a zeroed statistics buffer (forward) and a zeroed [2,C] sum buffer (backward), and every
# BatchNorm bumps `num_batches_tracked`: inside a `step_scope` these ~200 tiny fill / add launches collapse into one
# memset and one multi-tensor add.  The scope is opened INSIDE the region a CUDA graph captures, so replays redo both.
class _Scope:
    def __init__(self, device, capacity):
        self.buf = torch.zeros(capacity, dtype=torch.float32, device=device) if capacity else None
        self.used, self.asked, self.counters = 0, 0, []

    def take(self, n, device):
        n_al = (n + 31) // 32 * 32
        self.asked += n_al
        if self.buf is None or self.used + n_al > self.buf.numel() or self.buf.device != device:
            return torch.zeros(n, dtype=torch.float32, device=device)
        out = self.buf[self.used:self.used + n]
        self.used += n_al
        return out


_SCOPES = []
_SCOPE_CAPACITY = {}


class step_scope:
    """Context manager; `key` identifies the call site so that the pool is sized from the previous pass."""

    def __init__(self, key, device):
        self.key, self.device = (key, str(device)), device

    def __enter__(self):
        _SCOPES.append(_Scope(self.device, _SCOPE_CAPACITY.get(self.key, 0)))
        return self

    def __exit__(self, *exc):
        sc = _SCOPES.pop()
        _SCOPE_CAPACITY[self.key] = max(_SCOPE_CAPACITY.get(self.key, 0), sc.asked)
        if sc.counters and exc[0] is None:
            by_count = {}
            seen = {}
            for t in sc.counters:
                seen[id(t)] = (t, seen.get(id(t), (t, 0))[1] + 1)
            for t, c in seen.values():
                by_count.setdefault(c, []).append(t)
            for c, ts in by_count.items():
                torch._foreach_add_(ts, c)
        return False


# Backward passes run outside any scope (the autograd engine calls them).  `arm_backward_pool`, called at the end of a
# training forward, zeroes ONE pool (sized from the previous backward) that the backward's small zero-initialised buffers
# (atomic accumulators of bias / gamma / beta / token gradients, sparse weight gradients) are carved from, instead of one
# fill launch each (135 fills per C2 step before).  Slices are handed out once per arming, so they are always zero.
_BWD_POOL = {}


def arm_backward_pool(device):
    st = _BWD_POOL.setdefault(str(device), {"buf": None, "used": 0, "asked": 0})
    need = max(st["asked"], st["used"])
    if st["buf"] is None or st["buf"].numel() < need:
        st["buf"] = torch.zeros(int(need * 1.25) + 4096, dtype=torch.float32, device=device)
    elif st["used"]:
        st["buf"][:st["used"]].zero_()
    st["used"] = st["asked"] = 0


def zeros_f32(n, device):
    """Zeroed fp32 scratch of n elements (from the enclosing scope's pool, or the armed backward pool, when there is one)."""
    if _SCOPES:
        return _SCOPES[-1].take(n, device)
    st = _BWD_POOL.get(str(device))
    if st is not None and torch.cuda.is_current_stream_capturing():
        st = None       # a captured graph must own its fills: the pool is re-zeroed by eager code only
    if st is not None and st["buf"] is not None:
        n_al = (n + 31) // 32 * 32
        st["asked"] += n_al
        if st["used"] + n_al <= st["buf"].numel():
            out = st["buf"][st["used"]:st["used"] + n]
            st["used"] += n_al
            return out
    elif st is not None:
        st["asked"] += (n + 31) // 32 * 32
    return torch.zeros(n, dtype=torch.float32, device=device)


def bump_counter(t):
    """`num_batches_tracked += 1`, deferred to the end of the enclosing scope when there is one."""
    if _SCOPES:
        _SCOPES[-1].counters.append(t)
    else:
        t.add_(1)


def new_stats(co, device):
    return zeros_f32(STAT_COPIES * 2 * co, device).view(STAT_COPIES, 2, co)


# ================================================================================================ statistics exchange
# SyncBatchNorm-equivalent training (engine/train.py:160-161, `model.sync_bn: true` in both live configs): the batch
# statistics of every BatchNorm are taken over the frames / active sites of ALL ranks.  Per layer and direction ONE
# exchange of [2][C] sums (+ the element count in the forward): K15 (`mg_stats_exchange`) reduces the conv-epilogue
# copies, pushes the result into every rank's exchange window over peer memory (NVLink / NVSwitch), waits for the peers
# inside the kernel and sums in rank order - no NCCL launch, no host involvement, bit-identical statistics on all ranks.
# The global count stays on the device (`count_dev` of mg_bn_finalize / mg_bn_bwd_apply), so ranks with different numbers
# of active sites need no host synchronisation.  dgamma / dbeta stay LOCAL sums (the gradient all-reduce averages them,
# as DDP does).  Without peer windows (gloo groups, MAGGIE_B200_NO_PEER_EXCHANGE=1, IPC mapping refused) the same [2C+1]
# vector goes through one all-reduce of the process group instead.
_SYNC_ALL = None
_SYNC_EPOCH = 0
_WINDOWS = {}


def set_sync_bn(group=True):
    """Exchange the statistics of EVERY BatchNorm over `group` (True: the default group; None / False: off), without
    converting the containers.  Containers converted by `nn.SyncBatchNorm.convert_sync_batchnorm` exchange anyway."""
    global _SYNC_ALL, _SYNC_EPOCH
    _SYNC_ALL = group if group not in (False, None) else None
    _SYNC_EPOCH += 1


def sync_group(bn):
    """The process group `bn` shares its batch statistics over, or None (local statistics)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None
    if isinstance(bn, torch.nn.SyncBatchNorm):
        grp = bn.process_group if bn.process_group is not None else dist.group.WORLD
    elif _SYNC_ALL is not None:
        grp = dist.group.WORLD if _SYNC_ALL is True else _SYNC_ALL
    else:
        return None
    return grp if dist.get_world_size(grp) > 1 else None


def sync_bn_active(model):
    """True when some BatchNorm of `model` exchanges statistics."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2:
        return False
    return _SYNC_ALL is not None or any(isinstance(m, torch.nn.SyncBatchNorm) for m in model.modules())


def sync_bn_needs_eager(model, device):
    """The collective fallback of the exchange cannot be captured into the dense stage's CUDA graphs; K15 can."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2:
        return False
    key = (_SYNC_EPOCH, device.index)
    cached = model.__dict__.get("_mg_sync_eager")
    if cached is None or cached[0] != key:
        cached = model.__dict__["_mg_sync_eager"] = (key, _sync_bn_needs_eager(model, device))
    return cached[1]


def _sync_bn_needs_eager(model, device):
    if not sync_bn_active(model):
        return False
    import torch.distributed as dist
    groups = {id(g): g for g in (sync_group(m) for m in model.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm))
              if g is not None}
    if not groups:
        groups = {0: dist.group.WORLD}
    return any(peer_window(g, device) is None for g in groups.values())


def exchange(t, group):
    """In-place sum of `t` over `group`, ordered on the current stream (the collective fallback of K15).  A gloo group is
    served through a host copy of the few hundred floats."""
    import torch.distributed as dist
    if t.is_cuda and dist.get_backend(group) == "gloo":
        h = t.cpu()
        dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)
        t.copy_(h)
        return t
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class XchgDesc(ctypes.Structure):
    """Mirror of `mg_xchg_desc`."""
    _fields_ = [("window", c_void_p * 16), ("rank", c_int32), ("world", c_int32)]


class PeerWindow:
    """This rank's exchange window and the mapped windows of its peers (CUDA IPC; handles travel over the process group)."""

    def __init__(self, group, device):
        import torch.distributed as dist
        L = _lib.lib()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 16:
            raise RuntimeError("statistics exchange windows support up to 16 ranks")
        with torch.cuda.device(device):
            win, handle = c_void_p(), ctypes.create_string_buffer(64)
            mine = None
            if L.mg_xchg_window_create(ctypes.byref(win), handle) == 0:
                self.local, mine = win.value, bytes(handle.raw)
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=group)
            self.desc, self.mapped = XchgDesc(), []
            self.desc.rank, self.desc.world = self.rank, self.world
            ok = all(h is not None for h in handles)
            for r, h in enumerate(handles if ok else []):
                if r == self.rank:
                    self.desc.window[r] = self.local
                    continue
                p = c_void_p()
                if L.mg_xchg_window_open(ctypes.create_string_buffer(h, 64), ctypes.byref(p)) != 0:
                    ok = False
                    break
                self.mapped.append(p.value)
                self.desc.window[r] = p.value
            # everybody or nobody: one rank falling back to the collective alone would dead-lock the others
            votes = [None] * self.world
            dist.all_gather_object(votes, bool(ok), group=group)
            self.ok = all(votes)

    def ptr(self):
        return ctypes.byref(self.desc)


def peer_window(group, device):
    """The PeerWindow of (group, device), created on first use (a collective call); None: use the collective fallback."""
    key = (id(group), device.index)
    if key not in _WINDOWS:
        w = None
        import torch.distributed as dist
        if os.environ.get("MAGGIE_B200_NO_PEER_EXCHANGE", "0") != "1" and device.type == "cuda":
            try:
                w = PeerWindow(group, device)
                if not w.ok:
                    w = None
            except RuntimeError:
                w = None
        if w is None and dist.get_rank(group) == 0:
            import warnings
            warnings.warn("maggie_b200: BatchNorm statistics are exchanged with one all-reduce per layer "
                          "(peer-memory exchange windows unavailable or disabled)")
        _WINDOWS[key] = w
    return _WINDOWS[key]


def exchange_sums(src, n_copies, C, count, group):
    """[2][C] sums (`n_copies` partial copies) + element count (None: no count) of this rank -> sums over all ranks of
    `group`: fp32 [2*C + 1], the last element being the global count."""
    out = torch.empty(2 * C + 1, dtype=torch.float32, device=src.device)
    w = peer_window(group, src.device)
    cnt = -1.0 if count is None else float(count)
    _lib.check(_lib.lib().mg_stats_exchange(w.ptr() if w is not None else None, _ptr(src), n_copies, C, cnt, _ptr(out),
                                            _stream()), "mg_stats_exchange")
    if w is None:
        exchange(out if count is not None else out[:2 * C], group)
    return out


# ================================================================================================ wgrad (K4)
class WgradDesc(ctypes.Structure):
    """Mirror of `mg_wgrad_desc`."""
    _fields_ = [
        ("dy", c_void_p), ("N", c_int32), ("Hy", c_int32), ("Wy", c_int32), ("Co", c_int32),
        ("x", c_void_p), ("Hi", c_int32), ("Wi", c_int32), ("Ci", c_int32),
        ("dw", c_void_p), ("Ktot", c_int32),
        ("n_taps", c_int32), ("tap_dy", c_int32 * MAX_TAPS), ("tap_dx", c_int32 * MAX_TAPS), ("tap_koff", c_int32 * MAX_TAPS),
        ("sy", c_int32), ("sx", c_int32), ("ays", c_int32), ("ay0", c_int32), ("axs", c_int32), ("ax0", c_int32),
        ("Hg", c_int32), ("Wg", c_int32),
    ]


def wgrad_launch(dy, x, taps, dw, *, stride=1, dy_map=(1, 0, 1, 0), grid_hw):
    """dw fp32 [Co, Ktot] += sum_p dy[p mapped by dy_map] (x) x[p*stride + tap]  (see mg_conv_wgrad)."""
    _need_cuda(dy, x, dw)
    assert dy.dtype == torch.float16 and dy.is_contiguous() and x.dtype == torch.float16 and x.is_contiguous()
    assert dw.dtype == torch.float32 and dw.is_contiguous()
    d = WgradDesc()
    d.dy, d.N, d.Hy, d.Wy, d.Co = dy.data_ptr(), dy.shape[0], dy.shape[1], dy.shape[2], dy.shape[3]
    d.x, d.Hi, d.Wi, d.Ci = x.data_ptr(), x.shape[1], x.shape[2], x.shape[3]
    d.dw, d.Ktot = dw.data_ptr(), dw.shape[1]
    d.n_taps = len(taps)
    for i, (ty, tx, ko) in enumerate(taps):
        d.tap_dy[i], d.tap_dx[i], d.tap_koff[i] = ty, tx, ko
    d.sy = d.sx = stride
    d.ays, d.ay0, d.axs, d.ax0 = dy_map
    d.Hg, d.Wg = grid_hw
    _lib.check(_lib.lib().mg_conv_wgrad(ctypes.byref(d), _stream()), "mg_conv_wgrad")
    return dw


# ================================================================================================ auxiliary stream
# Weight gradients are leaves of the backward data-flow: nothing downstream of a layer waits for them until the grouped
# weight-prep backward at the very end.  They run on ONE auxiliary stream, forked after the layer's BatchNorm backward
# and joined once before `mg_wprep_bwd`, so that they overlap with the dgrad -> BN-backward chain of the layers below
# (most launches of this network fill only a fraction of the 148 SMs).  Fork and join are plain event waits, i.e.
# CUDA-graph capturable.
_AUX = {}
AUX_WGRAD = os.environ.get("MAGGIE_B200_NO_AUX_STREAM", "0") != "1"


def aux_stream(device, which=0):
    """which 0: the weight-gradient stream; 1: the encoder's shortcut-branch stream."""
    st = _AUX.get((device.index, which))
    if st is None:
        st = _AUX[(device.index, which)] = torch.cuda.Stream(device=device)
    return st


SIDE_BRANCHES = os.environ.get("MAGGIE_B200_NO_SIDE_SHORTCUTS", "0") != "1"


def side_branch(x, bn, which, fn):
    """fn() on auxiliary stream `which`, forked from the current stream -> (result, join).  `join()` must be called on the
    main stream before the result is used.  Runs inline (join = no-op) on the CPU, when disabled, and when `bn` exchanges
    statistics across ranks (all exchanges must then keep one stream order on every rank)."""
    if not (x.is_cuda and SIDE_BRANCHES and sync_group(bn) is None):
        return fn(), (lambda: None)
    main, side = torch.cuda.current_stream(x.device), aux_stream(x.device, which)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        y = fn()
    x.record_stream(side)

    def join():
        main.wait_stream(side)
        y.record_stream(main)

    return y, join


def wgrad_async(geom, dr, xn, w_shape, bank):
    """geom.wgrad into the bank's accumulator on the auxiliary stream (joined by WeightBank._backward)."""
    main = torch.cuda.current_stream(dr.device)
    side = aux_stream(dr.device)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        geom.wgrad(dr, xn, w_shape, bank=bank)
    dr.record_stream(side)
    xn.record_stream(side)
    bank.prep.aux = side


# ================================================================================================ geometry
class ConvGeom:
    """Forward / dgrad / wgrad launch recipes of one conv layer.  kind 'conv': k x k, stride 1|2, pad, dilation
    (torch weight [Co,Ci,k,k]); kind 'convT': ConvTranspose2d(4, 2, 1) (torch weight [Ci,Co,4,4])."""

    def __init__(self, kind="conv", k=3, stride=1, pad=1, dil=1):
        self.kind, self.k, self.stride, self.pad, self.dil = kind, k, stride, pad, dil
        assert kind in ("conv", "convT") and stride in (1, 2)
        assert not (stride == 2 and dil != 1)

    # ---- forward
    def out_hw(self, Hi, Wi):
        if self.kind == "convT":
            return 2 * Hi, 2 * Wi
        f = lambda n: (n + 2 * self.pad - self.dil * (self.k - 1) - 1) // self.stride + 1
        return f(Hi), f(Wi)

    def fwd(self, x, w, **epi):
        if self.kind == "convT":
            assert not epi or set(epi) <= {"stats"}
            return conv_transpose4x4s2_nhwc(x, w, **epi)
        return conv2d_nhwc(x, w, stride=self.stride, padding=self.pad, dilation=self.dil, **epi)

    # ---- data gradient: dx [N,Hi,Wi,Ci_pad] from dy [N,Ho,Wo,Co]
    def dgrad(self, dy, w, x_shape, res=None):
        """`res` (stride-1 convs only): NHWC fp16 tensor of the input's shape added in the epilogue (dx = conv^T(dy) + res)."""
        N, Hi, Wi, ci_pad = x_shape
        assert res is None or (self.kind == "conv" and self.stride == 1)
        k, p, dl = self.k, self.pad, self.dil
        if self.kind == "convT":
            Ci, Co = wshape(w)[:2]
            if _banked(w):
                wp = w.D
            else:
                wp = pack_weight(w, Co) if ci_pad == Ci else pack_weight(F.pad(w, (0, 0, 0, 0, 0, 0, 0, ci_pad - Ci)), Co)
            taps = [(ky - 1, kx - 1, (ky * 4 + kx) * Co) for ky in range(4) for kx in range(4)]
            return conv_launch(dy, wp, taps, stride=2, grid_hw=(Hi, Wi))
        Co, Ci = wshape(w)[:2]
        if _banked(w):
            assert ci_pad == w.ci_pad
            wp = w.D
        else:
            wt = w.permute(1, 0, 2, 3)  # [Ci,Co,k,k]
            if ci_pad != Ci:
                wt = F.pad(wt, (0, 0, 0, 0, 0, 0, 0, ci_pad - Ci))
            wp = pack_weight(wt, Co)      # [Ci_pad][k*k][Co]
        if self.stride == 1:
            taps = [(p - ky * dl, p - kx * dl, (ky * k + kx) * Co) for ky in range(k) for kx in range(k)]
            return conv_launch(dy, wp, taps, grid_hw=(Hi, Wi), res=res)
        assert Hi % 2 == 0 and Wi % 2 == 0
        dx = torch.empty((N, Hi, Wi, ci_pad), dtype=torch.float16, device=dy.device)
        phases = []
        for py in range(2):
            for px in range(2):
                taps = [((py + p - ky) // 2, (px + p - kx) // 2, (ky * k + kx) * Co)
                        for ky in range(k) if (py + p - ky) % 2 == 0 for kx in range(k) if (px + p - kx) % 2 == 0]
                phases.append((taps, py, px))
        if all(ph[0] for ph in phases) and sum(len(ph[0]) for ph in phases) <= MAX_TAPS:
            # the four output parities as ONE launch (blockIdx.z = phase)
            conv_launch(dy, wp, None, grid_hw=(Hi // 2, Wi // 2), out=dx, out_map=(2, 0, 2, 0), phases=phases)
            return dx
        for taps, py, px in phases:
            if not taps:
                dx[:, py::2, px::2] = 0
                continue
            conv_launch(dy, wp, taps, grid_hw=(Hi // 2, Wi // 2), out=dx, out_map=(2, py, 2, px))
        return dx

    # ---- weight gradient in the torch layout, fp32
    def wgrad(self, dy, x, w_shape, bank=None):
        """`bank` (a BankedWeight): accumulate into the bank's pre-zeroed fp32 buffer and return None."""
        ci_pad = x.shape[-1]
        k = self.k
        if self.kind == "convT":
            Ci, Co = w_shape[:2]
            dwp = bank.G if bank is not None else torch.zeros((Co, 16 * ci_pad), dtype=torch.float32, device=x.device)
            for py in range(2):
                for px in range(2):
                    taps = [((py + 1 - ky) // 2, (px + 1 - kx) // 2, (ky * 4 + kx) * ci_pad)
                            for ky in range(4) if (py + 1 - ky) % 2 == 0 for kx in range(4) if (px + 1 - kx) % 2 == 0]
                    wgrad_launch(dy, x, taps, dwp, dy_map=(2, py, 2, px), grid_hw=x.shape[1:3])
            return None if bank is not None else dwp.view(Co, 4, 4, ci_pad)[..., :Ci].permute(3, 0, 1, 2)
        Co, Ci = w_shape[:2]
        dwp = bank.G if bank is not None else torch.zeros((Co, k * k * ci_pad), dtype=torch.float32, device=x.device)
        wgrad_launch(dy, x, conv_taps(k, k, self.pad, self.dil, ci_pad), dwp, stride=self.stride, grid_hw=dy.shape[1:3])
        return None if bank is not None else dwp.view(Co, k, k, ci_pad)[..., :Ci].permute(0, 3, 1, 2)


# ================================================================================================ BN pieces (K3)
def bn_finalize(stats, count, bn, training, group=None):
    """-> scale, shift, mean, invstd (fp32 [C]) and, with `group`, the device scalar holding the global element count.
    Training: batch statistics from the conv epilogue + running-stat update; eval: running statistics.  `group`: the sums
    and the element count are first taken over all ranks of the process group (`exchange_sums`)."""
    C = bn.weight.shape[0]
    dev = bn.weight.device
    out = torch.empty((4, C), dtype=torch.float32, device=dev)
    L = _lib.lib()
    count_dev = None
    if training:
        bump_counter(bn.num_batches_tracked)
        if group is not None:
            stats = exchange_sums(stats, STAT_COPIES, C, count, group)
            count_dev = stats[2 * C:]
        _lib.check(L.mg_bn_finalize(_ptr(stats), float(count), _ptr(bn.weight), _ptr(bn.bias), _ptr(bn.running_mean),
                                    _ptr(bn.running_var), float(bn.momentum), float(bn.eps), _ptr(out[0]), _ptr(out[1]),
                                    _ptr(out[2]), _ptr(out[3]), C, _ptr(count_dev), _stream()), "mg_bn_finalize")
    else:
        _lib.check(L.mg_bn_finalize(None, 1.0, _ptr(bn.weight), _ptr(bn.bias), _ptr(bn.running_mean), _ptr(bn.running_var),
                                    0.0, float(bn.eps), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(out[3]), C, None,
                                    _stream()), "mg_bn_finalize")
    if group is not None:
        return out[0], out[1], out[2], out[3], count_dev
    return out[0], out[1], out[2], out[3]


class _ConvBNAct(torch.autograd.Function):
    """conv -> BN(batch stats) -> (+res) -> act   |   conv -> act -> BN   (act_first), training mode, all native."""

    @staticmethod
    def forward(ctx, x, w, gamma, beta, res, geom, bn, act, act_first, res_up, handle=None):
        """`handle` (BankedWeight): `w` is then the bank's token tensor (autograd link to the grouped weight prep)."""
        xn = x.permute(0, 2, 3, 1)
        assert xn.is_contiguous() and xn.dtype == torch.float16
        wd = handle if handle is not None else w.detach()
        w_shape = wshape(wd)
        Co = w_shape[1] if geom.kind == "convT" else w_shape[0]
        group = sync_group(bn)
        stats = new_stats(Co, x.device)
        r = geom.fwd(xn, wd, stats=stats) if not act_first else geom.fwd(xn, wd, stats=stats, pre_act=act)
        N, Ho, Wo, _ = r.shape
        y = torch.empty_like(r)
        rn = None
        if res is not None:
            rn = res.permute(0, 2, 3, 1)
            assert rn.is_contiguous() and rn.dtype == torch.float16
        if group is None and FUSED_BN_APPLY and Co <= 512:
            # local statistics: finalize + apply in ONE launch (every CTA adds up the statistic copies itself)
            bump_counter(bn.num_batches_tracked)
            out4 = torch.empty((4, Co), dtype=torch.float32, device=x.device)
            _lib.check(_lib.lib().mg_bn_train_apply(_ptr(stats), float(N * Ho * Wo), _ptr(bn.weight), _ptr(bn.bias),
                                                    _ptr(bn.running_mean), _ptr(bn.running_var), float(bn.momentum), float(bn.eps),
                                                    _ptr(out4), _ptr(r), _ptr(rn), int(res_up), _ptr(y), N, Ho, Wo, Co,
                                                    0 if act_first else ACT[act], _stream()), "mg_bn_train_apply")
            mean, invstd = out4[2], out4[3]
            ctx.sync = None
        else:
            scale, shift, mean, invstd, *cnt = bn_finalize(stats, N * Ho * Wo, bn, True, group)
            ctx.sync = (group, cnt[0]) if group is not None else None    # (group, global element count on the device)
            _lib.check(_lib.lib().mg_bn_apply(_ptr(r), _ptr(scale), _ptr(shift), _ptr(rn), int(res_up), _ptr(y), N, Ho, Wo, Co,
                                              0 if act_first else ACT[act], _stream()), "mg_bn_apply")
        ctx.save_for_backward(xn, None if handle is not None else wd, r, y, mean, invstd, gamma.detach())
        ctx.handle = handle
        ctx.sums = zeros_f32(2 * Co, x.device).view(2, Co) if _SCOPES else None  # zeroed now, filled by the backward
        ctx.cfg = (geom, act, act_first, res_up, res is not None, w_shape)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        xn, w, r, y, mean, invstd, gamma = ctx.saved_tensors
        geom, act, act_first, res_up, has_res, w_shape = ctx.cfg
        handle = ctx.handle
        if handle is not None:
            w = handle
        dy = gy.permute(0, 2, 3, 1)
        if not dy.is_contiguous() or dy.dtype != torch.float16:
            dy = dy.contiguous().to(torch.float16)
        N, Ho, Wo, Co = r.shape
        L = _lib.lib()
        sums, ctx.sums = ctx.sums, None
        if sums is None:
            sums = zeros_f32(2 * Co, r.device).view(2, Co)
        a_post = 0 if act_first else ACT[act]
        _lib.check(L.mg_bn_bwd_reduce(_ptr(dy), _ptr(y), _ptr(r), _ptr(mean), _ptr(invstd), _ptr(sums), N, Ho, Wo, Co, a_post,
                                      _stream()), "mg_bn_bwd_reduce")
        dr = torch.empty_like(r)
        dres = torch.empty_like(r) if has_res else None
        gsums, count_dev = sums, None
        if ctx.sync is not None:
            gsums, count_dev = exchange_sums(sums, 1, Co, None, ctx.sync[0]), ctx.sync[1]
        _lib.check(L.mg_bn_bwd_apply(_ptr(dy), _ptr(y), _ptr(r), _ptr(mean), _ptr(invstd), _ptr(gamma), _ptr(gsums), _ptr(dr),
                                     _ptr(dres), N, Ho, Wo, Co, a_post, ACT[act] if act_first else 0, _ptr(count_dev),
                                     _stream()), "mg_bn_bwd_apply")
        dw = None
        if ctx.needs_input_grad[1]:
            if handle is not None and AUX_WGRAD:
                wgrad_async(geom, dr, xn, w_shape, handle)
            else:
                dw = geom.wgrad(dr, xn, w_shape, bank=handle)
        dx = geom.dgrad(dr, w, xn.shape).permute(0, 3, 1, 2) if ctx.needs_input_grad[0] else None
        if dres is not None:
            dres = dres.permute(0, 3, 1, 2)
            if res_up:
                dres = F.avg_pool2d(dres.float(), 2).mul_(4.0).to(torch.float16)
        return dx, dw, sums[1], sums[0], dres, None, None, None, None, None, None


def conv_bn_act(x, w, bn, training, *, stride=1, padding=1, dilation=1, act="relu", act_first=False, residual=None,
                transposed=False, res_up=False):
    """x NCHW-shaped channels-last fp16 (channels padded to a multiple of 16); w in the reference layout, fp32.
    Training: conv(stats epilogue) + finalize + apply kernels with a native backward.  Eval: BatchNorm is folded
    into the conv epilogue (one kernel)."""
    banked = _banked(w)
    _need_cuda(x, None if banked else w)
    geom = ConvGeom("convT", 4, 2, 1, 1) if transposed else ConvGeom("conv", wshape(w)[-1], stride, padding, dilation)
    if x.dtype == torch.float32 and not training:
        return _conv_bn_act_x3(x, w, bn, geom, act, act_first, residual, res_up)
    if x.dtype != torch.float16 or not x.permute(0, 2, 3, 1).is_contiguous():
        x = x.to(torch.float16).contiguous(memory_format=torch.channels_last)
    if residual is not None and (residual.dtype != torch.float16 or not residual.permute(0, 2, 3, 1).is_contiguous()):
        residual = residual.to(torch.float16).contiguous(memory_format=torch.channels_last)
    if training:
        assert bn is not None
        if banked:
            return _ConvBNAct.apply(x, w.token, bn.weight, bn.bias, residual, geom, bn, act, act_first, res_up, w)
        return _ConvBNAct.apply(x, w, bn.weight, bn.bias, residual, geom, bn, act, act_first, res_up)
    xn = x.permute(0, 2, 3, 1)
    scale = shift = None
    if bn is not None:
        scale, shift, _, _ = bn_finalize(None, 1, bn, False)
    rn = residual.permute(0, 2, 3, 1) if residual is not None else None
    if geom.kind == "convT":
        # four phase launches write disjoint output pixels; epilogue fusion applies per launch
        N, Hi, Wi, ci_pad = xn.shape
        Ci, Co = wshape(w)[:2]
        y = torch.empty((N, 2 * Hi, 2 * Wi, Co), dtype=torch.float16, device=x.device)
        wp = w.P if banked else pack_weight(w.detach().permute(1, 0, 2, 3), ci_pad)
        conv_launch(xn, wp, None, grid_hw=(Hi, Wi), out=y, out_map=(2, 0, 2, 0), scale=scale, shift=shift,
                    pre_act=act if act_first else None, post_act=None if act_first else act, phases=convT_phases(ci_pad))
    else:
        y = conv2d_nhwc(xn, w if banked else w.detach(), stride=stride, padding=padding, dilation=dilation, scale=scale, shift=shift, res=rn,
                        res_up=res_up, pre_act=act if act_first else None, post_act=None if act_first else act)
    return y.permute(0, 3, 1, 2)


def _conv_bn_act_x3(x, w, bn, geom, act, act_first, residual, res_up):
    """Eval-mode conv (+ folded BatchNorm, residual, activation) at fp32-level accuracy: fp32 activations in and out,
    split-fp16 operands on the tensor cores (`precision="high"`, see MaGGIe.set_precision).  `w`: fp32 weight tensor."""
    if _banked(w):
        raise RuntimeError("the fp32-accurate evaluation mode takes fp32 weight tensors, not fp16 bank packs")
    if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad):
        raise RuntimeError("the fp32-accurate evaluation mode has no backward: run it under torch.no_grad()")
    xn = x.permute(0, 2, 3, 1).contiguous()
    ci_pad = xn.shape[-1]
    assert ci_pad % 16 == 0
    x_hi, x_lo = split_f32(xn)
    wd = w.detach()
    scale = shift = None
    if bn is not None:
        scale, shift, _, _ = bn_finalize(None, 1, bn, False)
    rn = residual.float().permute(0, 2, 3, 1).contiguous() if residual is not None else None
    pre, post = (act, None) if act_first else (None, act)
    if geom.kind == "convT":
        N, Hi, Wi, _ = xn.shape
        Co = wd.shape[1]
        w_hi, w_lo = split_f32(pack_weight_f32(wd.permute(1, 0, 2, 3), ci_pad))
        y = torch.empty((N, 2 * Hi, 2 * Wi, Co), dtype=torch.float32, device=x.device)
        conv_launch(x_hi, w_hi, None, grid_hw=(Hi, Wi), out=y, out_map=(2, 0, 2, 0), scale=scale, shift=shift, pre_act=pre,
                    post_act=post, phases=convT_phases(ci_pad), lo=(x_lo, w_lo))
    else:
        Co, _, kh, kw = wd.shape
        w_hi, w_lo = split_f32(pack_weight_f32(wd, ci_pad))
        Hi, Wi = xn.shape[1:3]
        Ho, Wo = geom.out_hw(Hi, Wi)
        y = conv_launch(x_hi, w_hi, conv_taps(kh, kw, geom.pad, geom.dil, ci_pad), stride=geom.stride, grid_hw=(Ho, Wo),
                        scale=scale, shift=shift, res=rn, res_up=res_up, pre_act=pre, post_act=post, lo=(x_lo, w_lo))
    return y.permute(0, 3, 1, 2)


def linear_rows_x3(x, w, b=None):
    """y = x W^T + b on fp32 rows [..., Cin] at fp32-level accuracy: the rows are viewed as an NHWC image and run through
    the 1x1 case of the split-operand tensor-core conv."""
    rows, cin = x.numel() // x.shape[-1], x.shape[-1]
    co = w.shape[0]
    assert cin % 16 == 0 and co % 16 == 0
    d = 16
    while rows % d:
        d //= 2
    xn = x.detach().float().reshape(1, rows // d, d, cin).contiguous()
    x_hi, x_lo = split_f32(xn)
    w_hi, w_lo = split_f32(w.detach().float().reshape(co, cin).contiguous())
    bias = b.detach().float().contiguous() if b is not None else None
    y = conv_launch(x_hi, w_hi, [(0, 0, 0)], bias=bias, lo=(x_lo, w_lo))
    return y.reshape(*x.shape[:-1], co)


class _ConvBias(torch.autograd.Function):
    """Plain conv (+ bias), no normalisation: ConvGRU gates, the 32->1 temporal-difference head."""

    @staticmethod
    def forward(ctx, x, w, bias, geom):
        xn = x.permute(0, 2, 3, 1)
        assert xn.is_contiguous() and xn.dtype == torch.float16
        wd = w.detach()
        b32 = bias.detach().float().contiguous() if bias is not None else None
        y = geom.fwd(xn, wd, bias=b32)
        ctx.save_for_backward(xn, wd)
        ctx.cfg = (geom, tuple(w.shape), bias is not None)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        xn, w = ctx.saved_tensors
        geom, w_shape, has_bias = ctx.cfg
        dy = gy.permute(0, 2, 3, 1)
        if not dy.is_contiguous() or dy.dtype != torch.float16:
            dy = dy.contiguous().to(torch.float16)
        dx = geom.dgrad(dy, w, xn.shape).permute(0, 3, 1, 2) if ctx.needs_input_grad[0] else None
        dw = geom.wgrad(dy, xn, w_shape) if ctx.needs_input_grad[1] else None
        db = dy.float().sum((0, 1, 2)) if has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db, None


def conv_bias(x, w, bias=None, *, padding=1):
    """x NCHW-shaped channels-last fp16; w [Co,Ci,k,k] fp32 (Co padded to a multiple of 16 by the caller)."""
    _need_cuda(x, w)
    if x.dtype != torch.float16 or not x.permute(0, 2, 3, 1).is_contiguous():
        x = x.to(torch.float16).contiguous(memory_format=torch.channels_last)
    return _ConvBias.apply(x, w, bias, ConvGeom("conv", w.shape[-1], 1, padding, 1))


# ================================================================================================ ConvGRU step (K11)
class _GRUStep(torch.autograd.Function):
    """One ConvGRU step (module/conv_gru.py:50-58), all native: concat -> conv_ih -> gate1 -> conv_hh -> gate2, and the
    reverse chain in the backward (the data gradient of conv_ih adds the directly propagated parts in its epilogue and
    returns [dx | dh] as one tensor).  x, h: NCHW-shaped channels-last fp16 [N, C, h, w]."""

    @staticmethod
    def forward(ctx, x, h, w_ih, b_ih, w_hh, b_hh):
        L = _lib.lib()
        xn, hn = x.permute(0, 2, 3, 1), h.permute(0, 2, 3, 1)
        assert xn.is_contiguous() and hn.is_contiguous() and xn.dtype == hn.dtype == torch.float16
        N, Hh, Ww, C = xn.shape
        P = N * Hh * Ww
        g = ConvGeom("conv", 3, 1, 1, 1)
        wi, wh = w_ih.detach(), w_hh.detach()
        cat1 = torch.empty((N, Hh, Ww, 2 * C), dtype=torch.float16, device=x.device)
        _lib.check(L.mg_gru_concat2(_ptr(xn), _ptr(hn), _ptr(cat1), P, C, _stream()), "mg_gru_concat2")
        rz = g.fwd(cat1, wi, bias=b_ih.detach().float().contiguous())
        cat2 = torch.empty_like(cat1)
        _lib.check(L.mg_gru_gate1_fwd(_ptr(rz), _ptr(cat1), _ptr(cat2), P, C, _stream()), "mg_gru_gate1_fwd")
        cpre = g.fwd(cat2, wh, bias=b_hh.detach().float().contiguous())
        hnew = torch.empty_like(xn)
        _lib.check(L.mg_gru_gate2_fwd(_ptr(rz), _ptr(cpre), _ptr(cat1), _ptr(hnew), P, C, _stream()), "mg_gru_gate2_fwd")
        ctx.save_for_backward(cat1, cat2, rz, cpre, wi, wh)
        return hnew.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gh):
        from . import ops
        L = _lib.lib()
        cat1, cat2, rz, cpre, wi, wh = ctx.saved_tensors
        N, Hh, Ww, C2 = cat1.shape
        C, P = C2 // 2, N * Hh * Ww
        g = ConvGeom("conv", 3, 1, 1, 1)
        dhn = gh.permute(0, 2, 3, 1)
        if not dhn.is_contiguous() or dhn.dtype != torch.float16:
            dhn = dhn.contiguous().to(torch.float16)
        drz, dcpre, dhd = torch.empty_like(rz), torch.empty_like(cpre), torch.empty_like(cpre)
        _lib.check(L.mg_gru_gate2_bwd(_ptr(dhn), _ptr(rz), _ptr(cpre), _ptr(cat1), _ptr(drz), _ptr(dcpre), _ptr(dhd), P, C,
                                      _stream()), "mg_gru_gate2_bwd")
        dcat2 = g.dgrad(dcpre, wh, cat2.shape)
        dpart = torch.empty_like(cat1)
        _lib.check(L.mg_gru_gate1_bwd(_ptr(dcat2), _ptr(rz), _ptr(cat1), _ptr(dhd), _ptr(drz), _ptr(dpart), P, C, _stream()),
                   "mg_gru_gate1_bwd")
        dcat1 = g.dgrad(drz, wi, cat1.shape, res=dpart)            # = [dx | dh]
        dx = dcat1[..., :C].permute(0, 3, 1, 2) if ctx.needs_input_grad[0] else None
        dh = dcat1[..., C:].permute(0, 3, 1, 2) if ctx.needs_input_grad[1] else None
        dwi = g.wgrad(drz, cat1, tuple(wi.shape)) if ctx.needs_input_grad[2] else None
        dbi = ops.col_sum(drz.view(P, C2)) if ctx.needs_input_grad[3] else None
        dwh = g.wgrad(dcpre, cat2, tuple(wh.shape)) if ctx.needs_input_grad[4] else None
        dbh = ops.col_sum(dcpre.view(P, C)) if ctx.needs_input_grad[5] else None
        return dx, dh, dwi, dbi, dwh, dbh


def gru_step(x, h, w_ih, b_ih, w_hh, b_hh):
    """x, h [N, C, h, w] channels-last fp16; w_ih [2C, 2C, 3, 3], w_hh [C, 2C, 3, 3] fp32 (reference layout)."""
    _need_cuda(x, h, w_ih, w_hh)
    cl = lambda t: t if (t.dtype == torch.float16 and t.permute(0, 2, 3, 1).is_contiguous()) else \
        t.to(torch.float16).contiguous(memory_format=torch.channels_last)
    return _GRUStep.apply(cl(x), cl(h), w_ih, b_ih, w_hh, b_hh)
