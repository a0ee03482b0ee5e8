"""Fused low-bit attention forward -- host side of csrc/attn.cu.

Mirrors the reference's kernel-level entry points:
  forward(q, k, v, q_scale, k_scale, tensor_layout, output_dtype, return_lse) -> (o, lse)
      non-causal: src/triton/attn_qk_int8_per_block.py:169-238
      causal    : src/triton/attn_qk_int8_per_block_causal.py (host `forward`, kernel `_attn_fwd_base`)
`lse` is the kernel's base-2 value log2(l)+m, [B,Hq,Nq] f32 (the API layer converts it, core.py:344-350).
"""
import torch

from . import _native as N
from . import _tensor as T


# INT4 K codes travel packed (two per byte) when the kernel unpacks them in shared memory; until that variant
# lands the INT4 entry points feed one code per int8 to the same kind::i8 contraction (identical arithmetic).
PACKED_K4_KERNEL = False


def _out_dtype(output_dtype, default):
    if output_dtype is None:
        return default
    name = str(output_dtype).split(".")[-1]
    return {"float16": torch.float16, "bfloat16": torch.bfloat16}.get(name, default)


def _forward(q, k, v, q_scale, k_scale, tensor_layout, output_dtype, return_lse, causal,
             qk_mode=N.QK_I8, pv_mode=N.PV_F16, compat_tail=False, v_scale=None, v_mean=None, kbits=None):
    if tensor_layout not in ("HND", "NHD"):
        raise ValueError(f"tensor_layout {tensor_layout} not supported")
    qt, kt, vt = T.as_torch(q), T.as_torch(k), T.as_torch(v)
    qs, ks = T.as_torch(q_scale), T.as_torch(k_scale)
    dev = T.require_cuda(qt, kt, vt, qs, ks)
    b, hq, nq, d, qsb, qsh, qsn = T.bhnd(qt, tensor_layout)
    _, hkv, nk, dk, ksb, ksh, ksn = T.bhnd(kt, tensor_layout)
    if pv_mode == N.PV_F16:
        _, _, _, _, vsb, vsh, vsn = T.bhnd(vt, tensor_layout)
        assert vt.dtype == torch.float16, "V must be float16 for the FP16 P.V path"
    else:
        vsb, vsh, vsn = vt.stride(0), vt.stride(1), vt.stride(2)  # [B,H,D,Npad]
    assert qt.dtype == torch.int8 and kt.dtype == torch.int8
    assert qs.dtype == torch.float32 and ks.dtype == torch.float32 and qs.is_contiguous() and ks.is_contiguous()
    if causal:
        assert nq == nk, "qo_len and kv_len must be equal for causal attention"
    odt = _out_dtype(output_dtype, torch.float16)
    o = torch.empty(qt.shape, dtype=odt, device=dev)
    _, _, _, _, osb, osh, osn = T.bhnd(o, tensor_layout)
    lse = torch.empty((b, hq, nq), dtype=torch.float32, device=dev) if return_lse else None
    flags = (N.ATTN_CAUSAL if causal else 0) | (N.ATTN_COMPAT_TAIL if compat_tail else 0)
    ptr = lambda t: None if t is None else T.as_torch(t).data_ptr()
    N.call("lowbit_attn_fwd", qt.data_ptr(), kt.data_ptr(), vt.data_ptr(), qs.data_ptr(), ks.data_ptr(),
           ptr(v_scale), ptr(v_mean), ptr(kbits), o.data_ptr(), ptr(lse),
           b, hq, hkv, nq, nk, d, qsb, qsh, qsn, ksb, ksh, ksn, vsb, vsh, vsn, osb, osh, osn,
           qk_mode, pv_mode, T.dtype_code(odt), flags, T.stream_ptr(dev))
    if lse is None:
        lse = torch.empty([0], dtype=torch.float32)  # the reference returns an empty CPU tensor (:204)
    return T.like(o, q), T.like(lse, q)


def forward(q, k, v, q_scale, k_scale, tensor_layout="HND", output_dtype=torch.float16, return_lse=False,
            compat_tail=False):
    """Non-causal INT8-QK / FP16-PV attention over pre-quantized codes."""
    return _forward(q, k, v, q_scale, k_scale, tensor_layout, output_dtype, return_lse, False, compat_tail=compat_tail)


def forward_causal(q, k, v, q_scale, k_scale, tensor_layout="HND", output_dtype=torch.float16, return_lse=False):
    """Causal INT8-QK / FP16-PV attention over pre-quantized codes (qo_len == kv_len)."""
    return _forward(q, k, v, q_scale, k_scale, tensor_layout, output_dtype, return_lse, True)
