"""Plug-and-play use as `scaled_dot_product_attention` -- the way the reference is dropped into CogVideoX:

    paddle.nn.functional.scaled_dot_product_attention = sageattn          example/sageattn_cogvideo.py:9-14
    F.scaled_dot_product_attention = sageattn_qk_int8_pv_fp16_triton      bench/video_test/sageattn_cogvideo_int8.py

The patched function is then called by the model with SDPA's own signature: positional (q, k, v) in [B, H, N, D],
keyword `is_causal`, and keywords the low-bit operator has no use for (`attn_mask`, `dropout_p`, `scale`,
`enable_gqa`), which the reference swallows in **kwargs (src/core.py:194-205).  `as_sdpa` builds that callable from any
entry point of this package; `patch_sdpa` is the monkey-patch as a context manager (torch, and Paddle when importable).
"""
import contextlib
from typing import Any, Callable, Optional


def as_sdpa(op: Optional[Callable[..., Any]] = None, strict: bool = False, **op_kwargs: Any):
    """-> f(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None, **kwargs) calling `op`
    (default: lowbit_fa_qk_int8_pv_fp16_triton) with tensor_layout="HND".  `scale` maps to `sm_scale`.  Like the
    reference, `attn_mask` and `dropout_p` are accepted and ignored; strict=True raises instead when a mask or a
    non-zero dropout is passed (an attention mask cannot be honoured by this kernel)."""
    from . import core
    op = op or core.lowbit_fa_qk_int8_pv_fp16_triton

    def sdpa(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None, **kwargs):
        if strict and (attn_mask is not None or dropout_p):
            raise ValueError("the low-bit attention operator takes no attn_mask / dropout_p")
        return op(query, key, value, tensor_layout="HND", is_causal=bool(is_causal), sm_scale=scale, **op_kwargs)

    sdpa.__name__ = "lowbit_fa_sdpa"
    return sdpa


@contextlib.contextmanager
def patch_sdpa(op: Optional[Callable[..., Any]] = None, **op_kwargs: Any):
    """with patch_sdpa(): model(...)  -- every scaled_dot_product_attention call inside runs the low-bit operator."""
    import torch.nn.functional as F
    f = as_sdpa(op, **op_kwargs)
    saved = [(F, F.scaled_dot_product_attention)]
    F.scaled_dot_product_attention = f
    try:
        import paddle  # noqa: F401  (absent in this image: the torch patch is what the tests exercise)
        saved.append((paddle.nn.functional, paddle.nn.functional.scaled_dot_product_attention))
        paddle.nn.functional.scaled_dot_product_attention = f
    except Exception:
        pass
    try:
        yield f
    finally:
        for mod, orig in saved:
            mod.scaled_dot_product_attention = orig
