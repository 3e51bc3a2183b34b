"""ctypes binding of libmaggie_b200.so (the C ABI declared in include/maggie_b200.h).

No fallback: if the shared object is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_size_t, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmaggie_b200.so")
_lib = None

_I32P = POINTER(c_int32)
_PP = POINTER(c_void_p)

# name -> (restype, argtypes); must list every symbol in include/maggie_b200.h
SIGNATURES = {
    "mg_version": (c_int, []),
    "mg_last_error": (c_char_p, []),
    "mg_launch_count": (c_ulonglong, []),
    "mg_reset_launch_count": (None, []),
    "mg_unknown_mask": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mg_unknown_mask_select": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p]),
    "mg_fuse_stage": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mg_transition_gt": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mg_gru_concat2": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "mg_gru_gate1_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "mg_gru_gate2_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "mg_gru_gate2_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "mg_gru_gate1_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "mg_temporal_fuse_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_size_t, c_void_p]),
    "mg_temporal_fuse_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                             c_size_t, c_void_p]),
    "mg_sites_workspace": (c_size_t, [c_int, c_int, c_int]),
    "mg_sites_count": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mg_sites_tables": (c_int, [c_void_p, c_int, c_int, c_int, _I32P, _PP, _PP, _PP, _PP, c_void_p]),
    "mg_mask_embed_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mg_conv_fprop": (c_int, [c_void_p, c_void_p]),
    "mg_conv_fprop_x3": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "mg_split_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mg_mask_embed_fwd_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mg_layer_norm_fwd_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_void_p]),
    "mg_token_logits_fwd_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mg_attn_tq_fwd_f32": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_void_p] * 6),
    "mg_attn_fq_fwd_f32": (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_void_p] * 2),
    "mg_conv_wgrad": (c_int, [c_void_p, c_void_p]),
    "mg_conv_halo_launches": (c_ulonglong, []),
    "mg_conv_mid_launches": (c_ulonglong, []),
    "mg_conv_midt_launches": (c_ulonglong, []),
    "mg_conv_midt_trace": (None, [c_void_p]),
    "mg_conv_mid_trace": (None, [c_void_p]),
    "mg_wgrad_halo_launches": (c_ulonglong, []),
    "mg_attn_tq_workspace_floats": (c_size_t, [c_int, c_int, c_int]),
    "mg_attn_tq_fwd": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_void_p] * 6),
    "mg_attn_tq_bwd": (c_int, [c_void_p] * 11 + [c_int] * 4 + [c_void_p] * 4),
    "mg_attn_fq_fwd": (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_void_p] * 2),
    "mg_attn_fq_bwd": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_void_p] * 4),
    "mg_loss_workspace_floats": (c_size_t, [c_int, c_int, c_int]),
    "mg_loss_fwd": (c_int, [c_void_p] * 8 + [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mg_loss_bwd": (c_int, [c_void_p] * 8 + [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p]),
    "mg_gather_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "mg_scatter_rows_add": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mg_sparse_conv": (c_int, [c_void_p, c_void_p]),
    "mg_sparse_wgrad": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "mg_bn_finalize": (c_int, [c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "mg_bn_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mg_bn_train_apply": (c_int, [c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p,
                                  c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mg_bn_bwd_reduce": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p]),
    "mg_bn_bwd_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mg_input_stage": (c_int, [c_void_p] * 8 + [c_int] * 5 + [c_void_p]),
    "mg_alpha_finalize": (c_int, [c_void_p, c_void_p] + [c_int] * 7 + [c_float, c_float, c_void_p]),
    "mg_xchg_window_bytes": (ctypes.c_size_t, []),
    "mg_xchg_window_create": (c_int, [c_void_p, c_void_p]),
    "mg_xchg_window_open": (c_int, [c_void_p, c_void_p]),
    "mg_xchg_window_close": (c_int, [c_void_p]),
    "mg_xchg_window_destroy": (c_int, [c_void_p]),
    "mg_stats_exchange": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    "mg_wprep_fwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int] + [c_void_p] * 5),
    "mg_wprep_bwd": (c_int, [c_void_p, c_void_p, c_int] + [c_void_p] * 5),
    "mg_optim_adamw_step": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
                            + [c_float] * 7 + [c_void_p, c_void_p]),
    "mg_upsample_tanh_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mg_upsample_tanh_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mg_layer_norm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                  c_void_p]),
    "mg_layer_norm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "mg_token_logits_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mg_token_logits_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mg_col_sum": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mg_mask_embed_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
}


def lib():
    """The loaded library; raises RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m maggie_b200._build` (or __graft_entry__.build()). "
                "maggie_b200 has no fallback path.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(code, what):
    if code != 0:
        raise RuntimeError(f"{what} failed ({code}): {lib().mg_last_error().decode()}")


def launch_count():
    return int(lib().mg_launch_count())


def reset_launch_count():
    lib().mg_reset_launch_count()


def i32_array(values):
    return (c_int32 * len(values))(*[int(v) for v in values])


def ptr_array(ptrs):
    return (c_void_p * len(ptrs))(*[c_void_p(p) if p else None for p in ptrs])


def stream_ptr():
    """Raw handle of torch's current CUDA stream (the C-level getters: `torch.cuda.current_stream()` costs ~20 us)."""
    import torch

    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


def tensor_ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def need_cuda(*tensors):
    """Every tensor must live on the CURRENT CUDA device: kernels are launched on that device's current stream
    (`stream_ptr`), so a tensor on another GPU would be touched from the wrong context / without stream ordering."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("maggie_b200 native op called with a non-CUDA tensor; there is no CPU fallback")
        if cur is None:
            import torch

            cur = torch._C._cuda_getDevice()
        if t.device.index != cur:
            raise RuntimeError(f"maggie_b200 native op called with a tensor on cuda:{t.device.index} while the current device "
                               f"is cuda:{cur}; call torch.cuda.set_device() (or use `with torch.cuda.device(...)`) first")
