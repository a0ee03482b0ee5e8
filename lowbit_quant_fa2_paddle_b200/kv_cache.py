"""Attention over a KIVI-packed low-bit K/V cache -- host side.

Mirrors `_quantized_flash_attn_forward` of the reference's KV-cache prototype
(`src/triton/quantization/attn_4bit_per_block.py:428-553`): fp16 queries `[B,Nq,H,D]` against a cache that
`triton_quantize_and_pack_along_last_dim` produced (`src/triton/utils/quant/new_pack.py:247-300`) -- K transposed to
`[B,D,H,N]` and packed along the sequence, V `[B,N,H,D]` packed along the channels, asymmetric codes with a scale and
a minimum per group of 32 (the prototype's driver, `attn_4bit_per_block.py:637-665`).  The kernel is the hand-written
sm_100a one in `csrc/kv_attn.cu`; there is no CPU fallback.
"""
import math

import torch

from . import _native as N
from . import _tensor as T
from .quant import triton_quantize_and_pack_along_last_dim


def quant_and_pack_kv(k, v, group_size: int = 32, bits: int = 4):
    """The prototype driver's cache preparation (`attn_4bit_per_block.py:655-661`): k, v `[B,N,H,D]` fp16 ->
    (k_code, k_scale, k_mn, v_code, v_scale, v_mn), K quantized per channel along the sequence (`k.transpose(1, 3)`),
    V per token along the channels."""
    kt, vt = T.as_torch(k), T.as_torch(v)
    k_code, k_scale, k_mn = triton_quantize_and_pack_along_last_dim(kt.transpose(1, 3), group_size=group_size, bit=bits)
    v_code, v_scale, v_mn = triton_quantize_and_pack_along_last_dim(vt, group_size=group_size, bit=bits)
    return k_code, k_scale, k_mn, v_code, v_scale, v_mn


def quantized_flash_attn_forward(q, kcode, kscale, kmn, vcode, vscale, vmn, group_size=None, bits=None, bias=None,
                                 causal=False, softmax_scale=None):
    """`_quantized_flash_attn_forward` (`attn_4bit_per_block.py:428-553`) -> (o, lse, softmax_scale).
    q `[B,Nq,H,D]` fp16; kcode `[B,D,H,N*bits/8]` int8 with kscale / kmn `[B,D,H,N/group]`; vcode `[B,N,H,D*bits/8]`
    int8 with vscale / vmn `[B,N,H,D/group]`.  o like q; lse `[B,H,ceil(Nq/128)*128]` f32 (`:493-496`), natural log,
    rows >= Nq zero.  bias and causal masking are not part of this path (the prototype's driver uses neither)."""
    if bias is not None:
        raise NotImplementedError("quantized_flash_attn_forward: bias is not supported")
    if causal:
        raise NotImplementedError("quantized_flash_attn_forward: causal masking is not supported on the KV-cache path")
    group_size = 32 if group_size is None else int(group_size)
    bits = 4 if bits is None else int(bits)
    qt = T.as_torch(q)
    tens = [T.as_torch(x) for x in (kcode, kscale, kmn, vcode, vscale, vmn)]
    dev = T.require_cuda(qt, *tens)
    kc, ks, km, vc, vs, vm = tens
    assert qt.dim() == 4 and qt.dtype == torch.float16, "Only support query to be fp16"
    B, Nq, H, D = qt.shape
    assert D <= 128, "FlashAttention only support head dimensions up to 128"
    if D not in (64, 128):
        raise ValueError(f"Unsupported head_dim: {D} (the kernel takes 64 or 128)")
    assert kc.dtype == vc.dtype == torch.int8, "Key and Value tensors must have the same type"
    Nk = vc.shape[1]
    assert tuple(kc.shape) == (B, D, H, Nk * bits // 8), "seqlen of quantized key is not correct"
    assert tuple(vc.shape) == (B, Nk, H, D * bits // 8), "head dimention of quantized value is not correct"
    assert Nk % group_size == 0
    for t, shp in ((ks, (B, D, H, Nk // group_size)), (km, (B, D, H, Nk // group_size)),
                   (vs, (B, Nk, H, D // group_size)), (vm, (B, Nk, H, D // group_size))):
        assert tuple(t.shape) == shp and t.dtype == torch.float16, "scale / minimum tensors do not match the codes"
    kc, ks, km, vc, vs, vm = (t.contiguous() for t in (kc, ks, km, vc, vs, vm))
    if qt.stride(-1) != 1:
        qt = qt.contiguous()
    softmax_scale = softmax_scale or 1.0 / math.sqrt(D)
    o = torch.empty_like(qt)
    nq_round = (Nq + 127) // 128 * 128
    lse = torch.zeros((B, H, nq_round), dtype=torch.float32, device=dev)
    ws = torch.empty(N.lib().lowbit_kv_attn_workspace_bytes(B, H, Nq, Nk, D), dtype=torch.uint8, device=dev)
    N.call("lowbit_kv_attn_fwd", qt.data_ptr(), kc.data_ptr(), ks.data_ptr(), km.data_ptr(), vc.data_ptr(),
           vs.data_ptr(), vm.data_ptr(), o.data_ptr(), lse.data_ptr(), ws.data_ptr(), B, H, Nq, Nk, D, group_size, bits,
           float(softmax_scale), qt.stride(0), qt.stride(1), qt.stride(2), o.stride(0), o.stride(1), o.stride(2),
           nq_round, T.stream_ptr(dev), device=dev)
    return T.like(o, q), T.like(lse, q), softmax_scale
