"""TEST INFRASTRUCTURE -- CPU oracle of the KV-cache path.  Only tests/ and measurement tools may
import this; the product path never does.

PARITY: the reference's `_fwd_kernel` (src/triton/quantization/attn_4bit_per_block.py:28-421) is a prototype that does
not run as written (`tl.arange` over a runtime length, `static_print`s, an int32-vs-int8 packing mismatch with its own
driver -- SURVEY 2.1 row 12), so no golden vector can come from the kernel itself.  It is pinned to what the
prototype's own driver checks the kernel against (`main`, :637-788: FlashAttention over `dequant_k`, `dequant_v`):
  * the cache format: `quantize_and_pack_along_last_dim` in oracle/quant.py is bit-exact against the reference's pack
    kernels (tests/golden/kivi_*.npz);
  * the dequantization: `kivi_unpack_and_dequant` in oracle/quant.py is bit-identical to the reference's
    `unpack_and_dequant_kcache` / `unpack_and_dequant_vcache` (new_pack.py:68-144), which tools/make_golden_kvcache.py
    executes UNMODIFIED through a Paddle shim (tests/golden/kvcache_*.npz: dequant_k, dequant_v);
  * the attention: exact softmax attention (fp64) over those dequantized fp16 caches is stored as o / lse in the same
    fixtures; this oracle is within 3.4e-4 / 1.1e-4 of it, the CUDA kernel within 5.4e-4 / 4.8e-4
    (tests/test_kv_cache.py).
On top of that format this file restates the arithmetic the prototype kernel states:
  K^ = fma(code, scale, mn), V^ = fma(code, scale, mn) in fp32            attn_4bit_per_block.py:260-262, 355-357
  S = q . K^ (fp32), p = exp(S * softmax_scale - m), o = sum p V^ / l     :330-372
  lse = m + log(l) (natural)                                              :372
(the fp32 fma differs from the reference's fp16 `code * scale + mn` by the two fp16 roundings, <= 3e-3 absolute).
"""
import math

import torch


def unpack_codes(code: torch.Tensor, bits: int) -> torch.Tensor:
    """int8 [..., T*bits/8] -> int32 codes [..., T]; code i of a byte sits at bits [i*bits, (i+1)*bits)
    (`_pack_along_last_dim`, new_pack.py:198-219: `element << i * bits`)."""
    per = 8 // bits
    b = code.to(torch.int32) & 0xFF
    parts = [(b >> (i * bits)) & ((1 << bits) - 1) for i in range(per)]
    return torch.stack(parts, dim=-1).reshape(*code.shape[:-1], code.shape[-1] * per)


def dequant_lastdim(code, scale, mn, group_size: int, bits: int) -> torch.Tensor:
    """fp32 `fma(code, scale, mn)` with one (scale, mn) per `group_size` codes of the last dim."""
    c = unpack_codes(code, bits).to(torch.float32)
    shp = c.shape
    c = c.reshape(*shp[:-1], shp[-1] // group_size, group_size)
    out = c * scale.to(torch.float32).unsqueeze(-1) + mn.to(torch.float32).unsqueeze(-1)  # exact product + one rounding ~ fma
    return out.reshape(shp)


def quantized_flash_attn_forward(q, kcode, kscale, kmn, vcode, vscale, vmn, group_size=32, bits=4, softmax_scale=None):
    """q [B,Nq,H,D] fp16; kcode [B,D,H,N*bits/8], kscale/kmn [B,D,H,N/G]; vcode [B,N,H,D*bits/8], vscale/vmn
    [B,N,H,D/G] -> (o fp16 [B,Nq,H,D], lse f32 [B,H,Nq], softmax_scale)."""
    B, Nq, H, D = q.shape
    softmax_scale = softmax_scale or 1.0 / math.sqrt(D)
    khat = dequant_lastdim(kcode, kscale, kmn, group_size, bits)   # [B,D,H,N]
    vhat = dequant_lastdim(vcode, vscale, vmn, group_size, bits)   # [B,N,H,D]
    s = torch.einsum("bqhd,bdhn->bhqn", q.to(torch.float32), khat) * softmax_scale
    m = s.amax(dim=-1, keepdim=True)
    p = torch.exp(s - m)
    l = p.sum(dim=-1, keepdim=True)
    o = torch.einsum("bhqn,bnhd->bqhd", p / l, vhat)
    lse = (m + torch.log(l)).squeeze(-1)
    return o.to(torch.float16), lse, softmax_scale
