"""Variable-length (packed, cu_seqlens) form of the INT8 operator -- drop-in for the reference's

  sageattn_varlen / lowbit_fa_varlen       src/core.py:356-491
  per_block_int8 (varlen)                  src/triton/quant_per_block_varlen.py:75-142
  forward (non-causal / causal varlen)     src/triton/attn_qk_int8_block_varlen.py:197-248,
                                           src/triton/attn_qk_int8_per_block_causal_varlen.py

Packed layout: q `[T_q, Hq, D]`, k / v `[T_k, Hkv, D]`, sequence i owns rows `cu_seqlens[i]:cu_seqlens[i+1]`.
Semantics kept from the reference: K is smoothed with ONE mean over all packed tokens (`k.mean(dim=0)`,
core.py:448 -- batch-global, unlike the padded operator); quantization blocks (128 Q rows / 64 K rows) restart at
every sequence; scales are packed block-major `[total_blocks, H]` with `cu_seqlens_*_scale` offsets
(quant_per_block_varlen.py:92-107); causal needs q_len == k_len per sequence.

Round-1 implementation: every sequence is one strided view `[1, n_i, H, D]` of the packed tensors and runs through
the same sm_100a kernels as the padded operator (TMA descriptors are built per sequence, rows past the sequence end
are zero-filled by TMA exactly like the reference's masked loads); the sequence lengths are read on the host once
per call.  A single-launch tile scheduler over cu_seqlens is the next step.
"""
from typing import Any, Optional

import torch

from . import _native as N
from . import _tensor as T
from . import attention as A
from . import quant as Qz


def _lens(cu):
    c = T.as_torch(cu)
    assert c.dim() == 1 and c.numel() >= 2 and c.is_contiguous(), "cu_seqlens must be a contiguous 1-D tensor"
    assert c.dtype in (torch.int32, torch.int64), "cu_seqlens must be int32 or int64"
    v = c.tolist()  # one device -> host read per call
    assert v[0] == 0 and all(b >= a for a, b in zip(v, v[1:])), "cu_seqlens must start at 0 and be non-decreasing"
    return v


def _cu_scale(cu, blk, device):
    nb = [(b - a + blk - 1) // blk for a, b in zip(cu, cu[1:])]
    out = [0]
    for n in nb:
        out.append(out[-1] + n)
    return out, torch.tensor(out, dtype=torch.int32, device=device)


def k_mean_varlen(k):
    """`k.mean(dim=0, keepdim=True)` of packed K `[T, H, D]` (core.py:448) with the exact-sum contract of k_mean."""
    kt = T.as_torch(k)
    return T.like(Qz.k_mean(kt.unsqueeze(0), "NHD").reshape(1, kt.shape[1], kt.shape[2]), k)


def per_block_int8_varlen(q, k, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, BLKQ=128, BLKK=64,
                          sm_scale=None, km=None, backend="triton"):
    """-> (q_int8 [T_q,Hq,D], q_scale [nblk_q,Hq] f32, k_int8, k_scale [nblk_k,Hkv], cu_seqlens_q_scale,
    cu_seqlens_k_scale), the reference's return tuple.  `km` ([1,Hkv,D], optional) fuses the reference's separate
    `k - km` pass into the K quantizer."""
    if backend not in Qz._MODES:
        raise ValueError(f"Unsupported quantization backend: {backend}")
    qt, kt = T.as_torch(q), T.as_torch(k)
    dev = T.require_cuda(qt, kt)
    assert qt.dim() == 3 and kt.dim() == 3, "packed tensors are [tokens, heads, head_dim]"
    cq, ck = _lens(cu_seqlens_q), _lens(cu_seqlens_k)
    assert len(cq) == len(ck), "cu_seqlens_q and cu_seqlens_k must describe the same batch"
    assert cq[-1] == qt.shape[0] and ck[-1] == kt.shape[0], "cu_seqlens[-1] must equal the packed token count"
    d = qt.shape[-1]
    if sm_scale is None:
        sm_scale = d ** -0.5
    mode = Qz._MODES[backend]
    outs = []
    for x, cu, blk, sm_arg, kmx in ((qt, cq, BLKQ, sm_scale * Qz.LOG2E, None), (kt, ck, BLKK, 1.0, km)):
        h = x.shape[1]
        cs, cs_t = _cu_scale(cu, blk, dev)
        codes = torch.empty(x.shape, dtype=torch.int8, device=dev)
        scale = torch.empty((cs[-1], h), dtype=torch.float32, device=dev)
        kmv = None if kmx is None else T.as_torch(kmx).reshape(1, h, d)
        for i in range(len(cu) - 1):
            a, b = cu[i], cu[i + 1]
            if b == a:
                continue
            sc_i = torch.empty((1, h, cs[i + 1] - cs[i]), dtype=torch.float32, device=dev)
            Qz._quant_one(x[a:b].unsqueeze(0), kmv, blk, 8, False, sm_arg, mode, "NHD",
                          out=(codes[a:b].unsqueeze(0), sc_i))
            scale[cs[i]:cs[i + 1]] = sc_i[0].t()
        outs += [codes, scale, cs_t]
    q_c, q_s, cqs, k_c, k_s, cks = outs
    return T.like(q_c, q), T.like(q_s, q), T.like(k_c, k), T.like(k_s, k), T.like(cqs, q), T.like(cks, k)


def forward_varlen(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, q_scale, k_scale, cu_seqlens_q_scale,
                   cu_seqlens_k_scale, output_dtype=torch.float16, causal=False, compat_tail=False):
    """Attention over packed, pre-quantized codes (attn_qk_int8_block_varlen.py:197-248 and its causal twin)."""
    qt, kt, vt = T.as_torch(q), T.as_torch(k), T.as_torch(v)
    qs, ks = T.as_torch(q_scale), T.as_torch(k_scale)
    dev = T.require_cuda(qt, kt, vt, qs, ks)
    cq, ck = _lens(cu_seqlens_q), _lens(cu_seqlens_k)
    cqs, cks = _lens(cu_seqlens_q_scale), _lens(cu_seqlens_k_scale)
    odt = A._out_dtype(output_dtype, torch.float16)
    o = torch.empty(qt.shape, dtype=odt, device=dev)
    for i in range(len(cq) - 1):
        a, b, c, e = cq[i], cq[i + 1], ck[i], ck[i + 1]
        if b == a:
            continue
        if e == c:  # no keys: the reference divides 0 by l = 1 (attn_qk_int8_block_varlen.py:168-189)
            o[a:b].zero_()
            continue
        qs_i = qs[cqs[i]:cqs[i + 1]].t().contiguous().unsqueeze(0)
        ks_i = ks[cks[i]:cks[i + 1]].t().contiguous().unsqueeze(0)
        A._forward(qt[a:b].unsqueeze(0), kt[c:e].unsqueeze(0), vt[c:e].unsqueeze(0), qs_i, ks_i, "NHD", odt, False,
                   causal, compat_tail=compat_tail, out=o[a:b].unsqueeze(0))
    return T.like(o, q)


def sageattn_varlen(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q: int, max_seqlen_k: int,
                    is_causal: bool = False, sm_scale: Optional[float] = None, smooth_k: bool = True, **kwargs: Any):
    """core.py:356-491.  q `[T_q,Hq,D]`, k / v `[T_k,Hkv,D]` fp16 / bf16, cu_seqlens int32 / int64 `[batch+1]`.
    Returns o `[T_q,Hq,D]` in q's dtype."""
    qt, kt, vt = T.as_torch(q), T.as_torch(k), T.as_torch(v)
    dtype = qt.dtype
    assert dtype in [torch.float16, torch.bfloat16], \
        "Input tensors must be in dtype of torch.float16 or torch.bfloat16"
    assert qt.device == kt.device == vt.device, "All tensors must be on the same device."
    assert qt.dtype == kt.dtype == vt.dtype, "All tensors must have the same dtype."
    head_dim_og = qt.shape[-1]
    if head_dim_og > 128:
        raise ValueError(f"Unsupported head_dim: {head_dim_og}")
    dev = T.require_cuda(qt, kt, vt)
    d_to = 64 if head_dim_og <= 64 else 128
    if head_dim_og != d_to:
        qt, kt, vt = (torch.nn.functional.pad(t, (0, d_to - head_dim_og)) for t in (qt, kt, vt))
    assert qt.stride(-1) == 1 and kt.stride(-1) == 1 and vt.stride(-1) == 1, "Last dim of qkv must be contiguous."
    backend = kwargs.get("quantization_backend", "triton")
    with torch.cuda.device(dev):
        if dtype == torch.bfloat16:
            vt = vt.to(torch.float16)
        km = k_mean_varlen(kt) if smooth_k else None
        if sm_scale is None:
            sm_scale = 1.0 / head_dim_og ** 0.5
        cq, ck = _lens(cu_seqlens_q), _lens(cu_seqlens_k)
        assert len(cq) == len(ck) and cq[-1] == qt.shape[0] and ck[-1] == kt.shape[0]
        mode = Qz._MODES[backend]
        hq, hkv = qt.shape[1], kt.shape[1]
        kmv = None if km is None else km.reshape(1, hkv, d_to)
        o = torch.empty(qt.shape, dtype=dtype, device=dev)
        for i in range(len(cq) - 1):
            a, b, c, e = cq[i], cq[i + 1], ck[i], ck[i + 1]
            if b == a:
                continue
            if e == c:
                o[a:b].zero_()
                continue
            if is_causal:
                assert b - a == e - c, "qo_len and kv_len must be equal for causal attention"
            k_c, k_s = Qz._quant_one(kt[c:e].unsqueeze(0), kmv, 64, 8, False, 1.0, mode, "NHD")
            q_c, q_s = Qz._quant_one(qt[a:b].unsqueeze(0), None, 128, 8, False, sm_scale * Qz.LOG2E, mode, "NHD")
            A._forward(q_c, k_c, vt[c:e].unsqueeze(0), q_s, k_s, "NHD", dtype, False, bool(is_causal),
                       compat_tail=bool(kwargs.get("compat_tail", False)), out=o[a:b].unsqueeze(0))
    return T.like(o[..., :head_dim_og], q)


lowbit_fa_varlen = sageattn_varlen
