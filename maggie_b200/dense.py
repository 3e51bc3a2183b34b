"""Native dense ops (K2 conv on tcgen05/TMA, ...) - python side of the C ABI: descriptor structs, weight packs,
tap tables.  Tensors are NHWC fp16 (`[N,H,W,C]` contiguous)."""
import ctypes
from ctypes import c_int32, c_void_p

import torch
import torch.nn.functional as F

from . import _lib
from .ops import _need_cuda, _ptr, _stream

MAX_TAPS = 16
STAT_COPIES = 64


class ConvDesc(ctypes.Structure):
    """Mirror of `mg_conv_desc` (include/maggie_b200.h)."""
    _fields_ = [
        ("x", c_void_p), ("N", c_int32), ("Hi", c_int32), ("Wi", c_int32), ("Ci", c_int32),
        ("w", c_void_p), ("Co", c_int32), ("Ktot", c_int32),
        ("n_taps", c_int32), ("tap_dy", c_int32 * MAX_TAPS), ("tap_dx", c_int32 * MAX_TAPS), ("tap_koff", c_int32 * MAX_TAPS),
        ("sy", c_int32), ("sx", c_int32), ("Hg", c_int32), ("Wg", c_int32),
        ("out", c_void_p), ("Ho", c_int32), ("Wo", c_int32), ("Cs", c_int32), ("c_off", c_int32),
        ("oys", c_int32), ("oy0", c_int32), ("oxs", c_int32), ("ox0", c_int32),
        ("epi_relu", c_int32), ("stats", c_void_p), ("bias", c_void_p),
    ]


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


def conv_taps(kh, kw, pad, dil, ci_pad):
    """Forward-conv tap table: [(dy, dx, koff)]."""
    return [(ky * dil - pad, kx * dil - pad, (ky * kw + kx) * ci_pad) for ky in range(kh) for kx in range(kw)]


def conv_launch(x, w_packed, taps, *, stride=1, grid_hw=None, out=None, out_hw=None, out_map=(1, 0, 1, 0), c_off=0,
                relu=False, stats=None, bias=None):
    """Generic launch of mg_conv_fprop.  x [N,Hi,Wi,Ci] fp16 NHWC; w_packed [Co,Ktot] fp16; taps [(dy,dx,koff)].
    grid_hw: logical output grid (defaults to ceil(Hi/stride)); out: preallocated NHWC fp16 (or None);
    out_map = (oys, oy0, oxs, ox0)."""
    _need_cuda(x, w_packed)
    assert x.dtype == torch.float16 and x.is_contiguous() and w_packed.dtype == torch.float16 and w_packed.is_contiguous()
    N, Hi, Wi, Ci = x.shape
    Co, Ktot = w_packed.shape
    Hg, Wg = grid_hw if grid_hw is not None else ((Hi + stride - 1) // stride, (Wi + stride - 1) // stride)
    if out is None:
        Ho, Wo = out_hw if out_hw is not None else (Hg, Wg)
        out = torch.empty((N, Ho, Wo, Co), dtype=torch.float16, device=x.device)
    assert out.dtype == torch.float16 and out.is_contiguous()
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
    d.epi_relu = int(relu)
    d.stats = stats.data_ptr() if stats is not None else None
    d.bias = bias.data_ptr() if bias is not None else None
    _lib.check(_lib.lib().mg_conv_fprop(ctypes.byref(d), _stream()), "mg_conv_fprop")
    return out


def conv2d_nhwc(x, w, *, stride=1, padding=1, dilation=1, relu=False, stats=None, bias=None, out=None, c_off=0):
    """Forward conv: x NHWC fp16 (channels already padded to a multiple of 16), w torch layout [Co,Ci,kh,kw]."""
    Co, Ci, kh, kw = w.shape
    ci_pad = x.shape[-1]
    wp = pack_weight(w, ci_pad)
    Hi, Wi = x.shape[1:3]
    Ho = (Hi + 2 * padding - dilation * (kh - 1) - 1) // stride + 1
    Wo = (Wi + 2 * padding - dilation * (kw - 1) - 1) // stride + 1
    return conv_launch(x, wp, conv_taps(kh, kw, padding, dilation, ci_pad), stride=stride, grid_hw=(Ho, Wo), relu=relu,
                       stats=stats, bias=bias, out=out, c_off=c_off)


def conv_transpose4x4s2_nhwc(x, w, *, stats=None):
    """ConvTranspose2d(k=4, s=2, p=1): w torch layout [Ci,Co,4,4]; four sub-pixel phase convs (2x2 taps each).
    out[2y+py, 2x+px] = sum_{ky,kx} w[:, :, ky, kx]^T x[y+dy, x+dx] with dy = (py + 1 - ky) / 2 (integer)."""
    Ci, Co, kh, kw = w.shape
    assert kh == 4 and kw == 4
    N, Hi, Wi, ci_pad = x.shape
    out = torch.empty((N, 2 * Hi, 2 * Wi, Co), dtype=torch.float16, device=x.device)
    wp = pack_weight(w.permute(1, 0, 2, 3), ci_pad)  # [Co][ky][kx][Ci]
    for py in range(2):
        for px in range(2):
            taps = []
            for ky in range(4):
                if (py + 1 - ky) % 2:
                    continue
                for kx in range(4):
                    if (px + 1 - kx) % 2:
                        continue
                    taps.append(((py + 1 - ky) // 2, (px + 1 - kx) // 2, (ky * 4 + kx) * ci_pad))
            conv_launch(x, wp, taps, grid_hw=(Hi, Wi), out=out, out_map=(2, py, 2, px), stats=stats)
    return out


def new_stats(co, device):
    return torch.zeros((STAT_COPIES, 2, co), dtype=torch.float32, device=device)
