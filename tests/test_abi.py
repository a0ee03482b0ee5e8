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
    assert L.lowbit_version() == 8
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


# the reference's public surface (src/__init__.py:1-17) and the leading parameters of each entry point
# (src/core.py:82-91, 194-205, 356-368, 495-508, 735-748, 945-956, 1064-1073)
REFERENCE_API = {
    "sageattn": ["q", "k", "v", "tensor_layout", "is_causal", "sm_scale", "return_lse"],
    "sageattn_varlen": ["q", "k", "v", "cu_seqlens_q", "cu_seqlens_k", "max_seqlen_q", "max_seqlen_k", "is_causal",
                        "sm_scale", "smooth_k"],
    "sageattn_qk_int8_pv_fp16_triton": ["q", "k", "v", "tensor_layout", "quantization_backend", "is_causal", "sm_scale",
                                        "smooth_k", "return_lse"],
    "sageattn_qk_int8_pv_fp16_cuda": ["q", "k", "v", "tensor_layout", "is_causal", "qk_quant_gran", "sm_scale",
                                      "pv_accum_dtype", "smooth_k", "smooth_v", "return_lse"],
    "sageattn_qk_int8_pv_fp8_cuda": ["q", "k", "v", "tensor_layout", "is_causal", "qk_quant_gran", "sm_scale",
                                     "pv_accum_dtype", "smooth_k", "smooth_v", "return_lse"],
    "sageattn_qk_int4_pv_fp16_triton": ["q", "k", "v", "tensor_layout", "quantization_backend", "is_causal", "sm_scale",
                                        "smooth_k", "return_lse"],
    "sageattn_multi_precision": ["q", "k", "v", "tensor_layout", "is_causal", "sm_scale", "return_lse"],
}
ALIASES = {"lowbit_fa_attn": "sageattn", "lowbit_fa_varlen": "sageattn_varlen",
           "lowbit_fa_multi_precision": "sageattn_multi_precision",
           "lowbit_fa_qk_int8_pv_fp16_triton": "sageattn_qk_int8_pv_fp16_triton",
           "lowbit_fa_qk_int8_pv_fp16_cuda": "sageattn_qk_int8_pv_fp16_cuda",
           "lowbit_fa_qk_int8_pv_fp8_cuda": "sageattn_qk_int8_pv_fp8_cuda",
           "lowbit_fa_qk_int4_pv_fp16_triton": "sageattn_qk_int4_pv_fp16_triton"}


def test_reference_public_surface():
    """Every name the reference package exports exists here with the same leading parameters (in order), accepts
    **kwargs like the reference, and the lowbit_fa_* names are the same objects as the sageattn_* ones."""
    import inspect

    import lowbit_quant_fa2_paddle_b200 as L
    for name, params in REFERENCE_API.items():
        fn = getattr(L, name)
        sig = inspect.signature(fn)
        got = [p.name for p in sig.parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD]
        assert got == params, f"{name}: {got}"
        assert any(p.kind == p.VAR_KEYWORD for p in sig.parameters.values()), f"{name} must accept **kwargs"
    for new, old in ALIASES.items():
        assert getattr(L, new) is getattr(L, old)


def test_more_error_conventions():
    import lowbit_quant_fa2_paddle_b200 as L
    from lowbit_quant_fa2_paddle_b200._native import LowbitNativeError
    h = torch.randn(1, 2, 64, 64).half()
    with pytest.raises(ValueError):
        L.sageattn_qk_int8_pv_fp16_cuda(h, h, h, qk_quant_gran="per_block")
    with pytest.raises(ValueError):
        L.sageattn_qk_int8_pv_fp16_cuda(h, h, h, pv_accum_dtype="fp64")
    with pytest.raises(ValueError):
        L.sageattn_qk_int8_pv_fp8_cuda(h, h, h, pv_accum_dtype="fp16")
    with pytest.raises(LowbitNativeError):
        L.sageattn(h, h, h, attn_mask=None, dropout_p=0.0)  # SDPA keywords tolerated; CPU tensors refused
    p = torch.randn(10, 2, 64).half()
    cu = torch.tensor([0, 4, 10], dtype=torch.int32)
    with pytest.raises(LowbitNativeError):
        L.lowbit_fa_varlen(p, p, p, cu, cu, 6, 6)
    with pytest.raises(AssertionError):
        L.lowbit_fa_varlen(p.float(), p.float(), p.float(), cu, cu, 6, 6)
    with pytest.raises(LowbitNativeError):
        L.lowbit_fa_host(h, h, h, device="cpu")
    # equal head groups, the last one cut once more (large + small) to shorten the exposed tail of the pipeline
    assert L.plan_chunks(4, 32, 32, "HND", None) == \
        [(b, h0, h0 + 16) for b in range(4) for h0 in (0, 16)][:-1] + [(3, 16, 28), (3, 28, 32)]
    for args in ((4, 32, 32, "HND", None), (2, 8, 2, "HND", 8), (1, 6, 6, "HND", 3), (3, 4, 4, "HND", 1), (2, 4, 4, "NHD", 8)):
        units = [(b, h) for b, h0, h1 in L.plan_chunks(*args) for h in range(h0, h1)]
        assert units == [(b, h) for b in range(args[0]) for h in range(args[2])]  # every (batch, kv head) once, in order
    assert L.plan_chunks(2, 8, 2, "NHD", 8) == [(0, 0, 2), (1, 0, 2)]
    with pytest.raises(ValueError):
        L.plan_chunks(1, 3, 2, "HND", 4)
