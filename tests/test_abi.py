"""CPU: the C-ABI library builds, loads, and exports every symbol include/lowbit_fa.h declares
(no compute calls without a GPU); the host layer fails loudly off-GPU instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


@pytest.fixture(scope="module")
def native():
    import __graft_entry__ as g
    g.build()
    from lowbit_quant_fa2_paddle_b200 import _native
    return _native


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "lowbit_fa.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lowbit_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(native):
    lib = ctypes.CDLL(native.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/lowbit_fa.h but not exported"
    # and the ctypes table binds exactly the header's surface
    assert sorted(native.SIGNATURES) == syms


def test_version_and_error_string(native):
    L = native.lib()
    assert L.lowbit_version() == 3
    assert isinstance(L.lowbit_last_error(), bytes)


def test_argument_validation_without_gpu(native):
    """Bad arguments are rejected before any CUDA call (so this runs on CPU)."""
    L = native.lib()
    rc = L.lowbit_quant_per_block(None, None, None, None, 1, 1, 1, 64, 0, 0, 0, 0, 0, 0, 64, 8, 0, 1.0, 0, 0, None)
    assert rc != 0 and b"null pointer" in L.lowbit_last_error()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = L.lowbit_quant_per_block(p, None, p, p, 1, 1, 1, 96, 96, 96, 96, 96, 96, 96, 64, 8, 0, 1.0, 0, 0, None)
    assert rc != 0 and b"head_dim" in L.lowbit_last_error()
    rc = L.lowbit_quant_per_block(p, None, p, p, 1, 1, 1, 64, 64, 64, 64, 64, 64, 64, 64, 3, 0, 1.0, 0, 0, None)
    assert rc != 0 and b"bits" in L.lowbit_last_error()
    with pytest.raises(native.LowbitNativeError):
        native.call("lowbit_attn_fwd", p, p, p, p, p, None, None, None, p, None, 1, 3, 2, 8, 8, 64,
                    *([64] * 12), 0, 0, 0, 0, None)


def test_no_cpu_fallback():
    """CPU tensors are refused: the product path never routes through the oracle or torch math."""
    import lowbit_quant_fa2_paddle_b200 as L
    from lowbit_quant_fa2_paddle_b200._native import LowbitNativeError
    q = torch.randn(1, 2, 64, 64).half()
    with pytest.raises(LowbitNativeError):
        L.lowbit_fa_qk_int8_pv_fp16_triton(q, q, q)
    with pytest.raises(LowbitNativeError):
        L.per_block_int8(q, q)


def test_api_error_conventions():
    """Same exception types as the reference (core.py:269-275,287; quant_per_block.py:203)."""
    import lowbit_quant_fa2_paddle_b200 as L
    q = torch.randn(1, 2, 64, 64)
    with pytest.raises(AssertionError):
        L.lowbit_fa_qk_int8_pv_fp16_triton(q, q, q)  # fp32 not allowed
    h = q.half()
    with pytest.raises(AssertionError):
        L.lowbit_fa_qk_int8_pv_fp16_triton(h, h.bfloat16(), h)  # mixed dtypes
    with pytest.raises(ValueError):
        L.lowbit_fa_qk_int8_pv_fp16_triton(h, h, h, tensor_layout="BSHD")
    with pytest.raises(ValueError):
        L.lowbit_fa_qk_int8_pv_fp16_triton(h, h, h, quantization_backend="numpy")
    big = torch.randn(1, 1, 8, 256).half()
    with pytest.raises(ValueError):
        L.lowbit_fa_qk_int8_pv_fp16_triton(big, big, big)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lowbit_quant_fa2_paddle_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\.*oracle", txt, flags=re.M), f"{f} imports the oracle"
                assert "oracle/" not in txt and "oracle." not in txt, f"{f} references the oracle package"
