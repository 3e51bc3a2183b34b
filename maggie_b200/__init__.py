"""maggie_b200: B200-native (sm_100a) implementation of the MaGGIe forward/backward hot path.

`maggie_b200.network` mirrors the reference's `maggie.network` API (build_model, MaGGIe, state-dict names);
the kernels live in `csrc/` behind the C ABI of `include/maggie_b200.h` (libmaggie_b200.so, bound with
ctypes in `_lib.py`).  There is no CPU or library fallback for the native ops.
"""
__version__ = "0.1.0"
