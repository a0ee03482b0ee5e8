"""ctypes binding of liblowbit_fa_b200.so (the C ABI declared in include/lowbit_fa.h).

There is no CPU fallback and no alternative backend: if the shared library is missing, or a call
fails, this raises.  Build with `python -c "import __graft_entry__ as g; g.build()"` (or `make -C
lowbit_quant_fa2_paddle_b200/csrc`).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LOWBIT_LIB") or os.path.join(_HERE, "liblowbit_fa_b200.so")  # LOWBIT_LIB: A/B builds (tools/)

F16, BF16 = 0, 1
QMODE_TRITON, QMODE_CUDA = 0, 1
QMODE_FLAG_DIV_FULL = 0x200  # Q1 with PTX div.full.f32: the reference kernels as Triton JIT-compiles them for a GPU
QMODE_FLAG_IEEE_DIV = 0x100  # validation: Q1 quotient by IEEE division instead of the 3-instruction exact sequence
QK_I8, QK_Q8K4, QK_Q8KMIX, QK_F16 = 0, 1, 2, 3
PV_F16, PV_E4M3 = 0, 1
ATTN_CAUSAL, ATTN_COMPAT_TAIL, ATTN_NARROW = 1, 2, 4

_c = ctypes
_P, _I, _L, _F = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float

# name -> (restype, argtypes); must list every symbol of include/lowbit_fa.h (tests/test_abi.py checks)
SIGNATURES = {
    "lowbit_version": (_I, []),
    "lowbit_last_error": (_c.c_char_p, []),
    "lowbit_quant_per_block_varlen": (_I, [_P] * 6 + [_I] * 4 + [_L] * 4 + [_I] * 4 + [_F, _I, _I, _P]),
    "lowbit_attn_fwd_varlen": (_I, [_P] * 11 + [_I] * 7 + [_L] * 8 + [_I] * 5 + [_P]),
    "lowbit_quant_k_mixed": (_I, [_P] * 6 + [_I] * 4 + [_L] * 6 + [_F, _F, _I, _I, _P]),
    "lowbit_sub_mean": (_I, [_P, _P, _P] + [_I] * 4 + [_L] * 6 + [_I, _P]),
    "lowbit_k_mean_workspace_bytes": (_L, [_I] * 4),
    "lowbit_k_mean": (_I, [_P, _P, _P] + [_I] * 4 + [_L] * 3 + [_I, _P]),
    "lowbit_k_smooth_quant_supported": (_I, [_I, _I, _I]),
    "lowbit_k_smooth_quant": (_I, [_P, _P, _P, _P] + [_I] * 4 + [_L] * 6 + [_I, _I, _I, _I, _P]),
    "lowbit_quant_per_block": (_I, [_P, _P, _P, _P] + [_I] * 4 + [_L] * 6 + [_I, _I, _I, _F, _I, _I, _P]),
    "lowbit_quant_per_thread": (_I, [_P, _P, _P, _P] + [_I] * 4 + [_L] * 6 + [_I] * 5 + [_P]),
    "lowbit_quant_pack_lastdim": (_I, [_P, _P, _P, _P, _L, _I, _I, _I, _I, _P]),
    "lowbit_kv_attn_workspace_bytes": (_L, [_I] * 5),
    "lowbit_kv_attn_fwd": (_I, [_P] * 10 + [_I] * 7 + [_F] + [_L] * 6 + [_I, _P]),
    "lowbit_v_fp8_workspace_bytes": (_L, [_I] * 4),
    "lowbit_v_fp8_per_channel": (_I, [_P, _P, _P, _P, _P] + [_I] * 4 + [_L] * 6 + [_F, _I, _P]),
    "lowbit_abs_max": (_I, [_P, _P] + [_I] * 4 + [_L] * 3 + [_I, _P]),
    "lowbit_min_max": (_I, [_P, _P] + [_I] * 4 + [_L] * 3 + [_I, _P]),
    "lowbit_attn_fwd": (_I, [_P] * 10 + [_I] * 6 + [_L] * 12 + [_I] * 4 + [_P]),
    "lowbit_attn_fwd_partial": (_I, [_P] * 11 + [_I] * 6 + [_L] * 9 + [_L, _L] + [_I] * 4 + [_P]),
    "lowbit_attn_finalize": (_I, [_P] * 5 + [_I] * 4 + [_L] * 3 + [_I, _P]),
    "lowbit_attn_set_debug_buffer": (None, [_P]),
    "lowbit_lse_fixup": (_I, [_P, _P, _P] + [_I] * 5 + [_L] * 3 + [_F, _I, _P]),
    "lowbit_fa_fwd_workspace_bytes": (_L, [_I] * 8),
    "lowbit_fa_fwd": (_I, [_P] * 6 + [_I] * 7 + [_L] * 12 + [_F, _F] + [_I] * 7 + [_P]),
}

_lib = None


class LowbitNativeError(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LowbitNativeError(
                f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built "
                "(run __graft_entry__.build()); there is no CPU or Triton fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.lowbit_version() != 8:
            raise LowbitNativeError("liblowbit_fa_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def _current_device():
    import torch
    return torch.cuda.current_device()


def call(name, *args, device=None):
    """Call an int-returning entry point; raise with the library's error text on failure.  `device`: the CUDA device
    the tensors (and the stream argument) belong to -- made current for the call, so that a process that drives
    several GPUs launches on the right one whatever torch's current device is."""
    L = lib()
    if device is not None and device.index is not None and device.index != _current_device():
        import torch
        with torch.cuda.device(device):
            rc = getattr(L, name)(*args)
    else:
        rc = getattr(L, name)(*args)
    if rc != 0:
        raise LowbitNativeError(f"{name} failed: {L.lowbit_last_error().decode()}")
