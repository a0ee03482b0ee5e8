"""lowbit_quant_fa2_paddle_b200 -- B200-native low-bit FlashAttention operator behind the API of
Charles2530/lowbit_quant_fa2_paddle (`src/__init__.py:1-17`): per-block INT8/INT4 quantizers with K
mean-smoothing and a fused tcgen05/TMEM/TMA attention kernel, hand-written for sm_100a."""
from .core import (  # noqa: F401
    # legacy names (backward compatible)
    sageattn,
    sageattn_qk_int8_pv_fp16_cuda,
    sageattn_qk_int8_pv_fp16_triton,
    sageattn_qk_int4_pv_fp16_triton,
    sageattn_multi_precision,
    sageattn_qk_int8_pv_fp8_cuda,
    # preferred names
    lowbit_fa_attn,
    lowbit_fa_qk_int8_pv_fp16_cuda,
    lowbit_fa_multi_precision,
    lowbit_fa_qk_int8_pv_fp16_triton,
    lowbit_fa_qk_int4_pv_fp16_triton,
    lowbit_fa_q_int8_k_int4_pv_fp16,
    lowbit_fa_q_int8_k_dynamic,
    lowbit_fa_qk_int8_pv_fp8_cuda,
    lowbit_fa_qk_int4_pv_fp8,
    lowbit_fa_fp16,
    compute_scale,
    select_quantization,
)
from .quant import (  # noqa: F401
    k_mean,
    k_smooth_quant,
    k_smooth_quant_supported,
    per_block_int8,
    per_block_int8_cuda,
    per_block_int4_unpack,
    per_block_int4,
    per_block_q_int8_k_int4,
    per_block_k_lowbit,
    per_block_k_mixed,
    per_thread_int8,
    per_thread_int4,
    per_warp_int8,
    per_channel_fp8,
    sub_mean,
    sub_mean_given,
    triton_quantize_and_pack_along_last_dim,
)
from .varlen import (  # noqa: F401
    sageattn_varlen,
    lowbit_fa_varlen,
    per_block_int8_varlen,
    forward_varlen,
    k_mean_varlen,
)
from .host import lowbit_fa_host, plan_chunks  # noqa: F401
from .plugin import as_sdpa, patch_sdpa  # noqa: F401
from .kv_cache import quantized_flash_attn_forward, quant_and_pack_kv  # noqa: F401
from .attention import forward, forward_causal, forward_partial, finalize, PartialState  # noqa: F401

# Every public callable runs on the caller's stream: torch's current stream, or -- for Paddle tensors -- Paddle's
# (see _tensor.on_callers_stream).  One wrapper per function object, installed under every name that refers to it (here
# and in the defining module), so aliases stay identical objects.
import sys as _sys  # noqa: E402
import types as _types  # noqa: E402

from . import _tensor as _T  # noqa: E402


def _install_stream_wrappers():
    skip = {"plan_chunks", "as_sdpa", "patch_sdpa"}
    pkg = _sys.modules[__name__]
    wrapped = {}
    for name, obj in list(vars(pkg).items()):
        if isinstance(obj, _types.FunctionType) and not name.startswith("_") and name not in skip:
            if id(obj) not in wrapped:
                wrapped[id(obj)] = _T.on_callers_stream(obj)
    mods = [pkg] + [m for n, m in list(_sys.modules.items()) if n.startswith(__name__ + ".") and m is not None]
    for m in mods:
        for name, obj in list(vars(m).items()):
            if isinstance(obj, _types.FunctionType) and id(obj) in wrapped:
                setattr(m, name, wrapped[id(obj)])


_install_stream_wrappers()

__version__ = "0.1.0"
