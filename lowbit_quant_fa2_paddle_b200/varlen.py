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

Like the reference's kernels, nothing here reads the sequence lengths on the host: the grids are sized from
`max_seqlen_*` (grid = (ceil(max_seqlen / block), heads, sequences), quant_per_block_varlen.py:120 and
attn_qk_int8_block_varlen.py:126-128), a CTA whose tile lies past its sequence's end exits, and the per-sequence scale
offsets are a device-side cumsum.  One K-mean call, two quantizer launches and one attention launch serve the whole
batch.  Internally scales are head-major `[H, cap]` (cap = T // block + sequences, an upper bound that needs no
device read); `per_block_int8_varlen` / `forward_varlen` convert to and from the reference's `[blocks, H]` layout (that
conversion needs the exact block count, i.e. one device read, exactly where the reference has one too:
`paddle.empty((cu_seqlens_q_scale[-1], h_qo))`, quant_per_block_varlen.py:101-106).
"""
from typing import Any, Optional

import torch

from . import _native as N
from . import _tensor as T
from . import attention as A
from . import quant as Qz


def _cu32(cu, dev):
    c = T.as_torch(cu)
    assert c.dim() == 1 and c.numel() >= 2 and c.is_contiguous(), "cu_seqlens must be a contiguous 1-D tensor"
    assert c.dtype in (torch.int32, torch.int64), "cu_seqlens must be int32 or int64"
    return c.to(device=dev, dtype=torch.int32)


def _cu_scale(cu32, blk):
    """Device-side offsets of every sequence's first scale block: [0, cumsum(ceil(len / blk))], int32 (no host read)."""
    nb = (cu32[1:] - cu32[:-1] + (blk - 1)) // blk
    return torch.cat([torch.zeros(1, dtype=torch.int32, device=cu32.device), torch.cumsum(nb, 0).to(torch.int32)])


def _quant_packed(x, km, cu32, max_seqlen, blk, bits, pack, sm_arg, mode):
    """Per-sequence per-block quantization of packed x [T,H,D] in one launch.
    -> (codes [T,H,D*bits/8 if pack], scale head-major [H, cap], cu_scale int32 [nseq+1], cap)."""
    t, h, d = x.shape
    dev = x.device
    nseq = cu32.numel() - 1
    cap = t // blk + nseq
    dd = d * bits // 8 if (pack and bits < 8) else d
    codes = torch.empty((t, h, dd), dtype=torch.int8, device=dev)
    scale = torch.zeros((h, max(cap, 1)), dtype=torch.float32, device=dev)
    cs = _cu_scale(cu32, blk)
    if t > 0:
        N.call("lowbit_quant_per_block_varlen", x.data_ptr(), km.data_ptr() if km is not None else None,
               codes.data_ptr(), scale.data_ptr(), cu32.data_ptr(), cs.data_ptr(), nseq, h, int(max_seqlen), d,
               x.stride(1), x.stride(0), codes.stride(1), codes.stride(0), scale.stride(0), blk, bits,
               int(bool(pack)), float(sm_arg), mode, T.dtype_code(x.dtype), T.stream_ptr(dev), device=dev)
    return codes, scale, cs, cap


def _attend_packed(q_c, k_c, v, q_s, k_s, cu_q, cu_k, cqs, cks, max_seqlen_q, out_dtype, causal, qk_mode=N.QK_I8,
                   kbits=None):
    tq, hq, d = q_c.shape
    tk, hkv = k_c.shape[0], k_c.shape[1]
    dev = q_c.device
    o = torch.empty((tq, hq, d), dtype=out_dtype, device=dev)
    if tq > 0:
        N.call("lowbit_attn_fwd_varlen", q_c.data_ptr(), k_c.data_ptr(), v.data_ptr(), q_s.data_ptr(), k_s.data_ptr(),
               kbits.data_ptr() if kbits is not None else None, cu_q.data_ptr(), cu_k.data_ptr(), cqs.data_ptr(),
               cks.data_ptr(), o.data_ptr(), cu_q.numel() - 1, hq, hkv, tq, tk, int(max_seqlen_q), d,
               q_c.stride(1), q_c.stride(0), k_c.stride(1), k_c.stride(0), v.stride(1), v.stride(0), o.stride(1),
               o.stride(0), q_s.stride(0), k_s.stride(0), qk_mode, T.dtype_code(out_dtype),
               N.ATTN_CAUSAL if causal else 0, T.stream_ptr(dev), device=dev)
    return o


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
    assert qt.stride(-1) == 1 and kt.stride(-1) == 1, "Last dim of qkv must be contiguous."
    cq, ck = _cu32(cu_seqlens_q, dev), _cu32(cu_seqlens_k, dev)
    assert cq.numel() == ck.numel(), "cu_seqlens_q and cu_seqlens_k must describe the same batch"
    d = qt.shape[-1]
    if d not in (64, 128):
        raise ValueError(f"Unsupported head_dim: {d} (the kernels take 64 or 128; lowbit_fa_varlen pads smaller ones)")
    if sm_scale is None:
        sm_scale = d ** -0.5
    mode = Qz._MODES[backend]
    kmv = None if km is None else T.as_torch(km).reshape(kt.shape[1], d).contiguous()
    with torch.cuda.device(dev):
        q_c, q_hm, cqs, _ = _quant_packed(qt, None, cq, max_seqlen_q, BLKQ, 8, False, sm_scale * Qz.LOG2E, mode)
        k_c, k_hm, cks, _ = _quant_packed(kt, kmv, ck, max_seqlen_k, BLKK, 8, False, 1.0, mode)
        nq, nk = int(cqs[-1]), int(cks[-1])  # the reference reads these too (quant_per_block_varlen.py:101-106)
        q_s, k_s = q_hm[:, :nq].t().contiguous(), k_hm[:, :nk].t().contiguous()
    return T.like(q_c, q), T.like(q_s, q), T.like(k_c, k), T.like(k_s, k), T.like(cqs, q), T.like(cks, k)


def forward_varlen(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, q_scale, k_scale, cu_seqlens_q_scale,
                   cu_seqlens_k_scale, output_dtype=torch.float16, causal=False, compat_tail=False):
    """Attention over packed, pre-quantized codes (attn_qk_int8_block_varlen.py:197-248 and its causal twin); scales
    in the reference's `[blocks, H]` layout.  compat_tail=True (the reference's unmasked tail keys) is
    only defined sequence by sequence -- the rows after a packed sequence are the next sequence's -- and takes the
    per-sequence path; the default masks tail keys and runs as one launch."""
    qt, kt, vt = T.as_torch(q), T.as_torch(k), T.as_torch(v)
    qs, ks = T.as_torch(q_scale), T.as_torch(k_scale)
    dev = T.require_cuda(qt, kt, vt, qs, ks)
    odt = A._out_dtype(output_dtype, torch.float16)
    with torch.cuda.device(dev):
        if compat_tail:
            return T.like(_forward_varlen_per_sequence(qt, kt, vt, cu_seqlens_q, cu_seqlens_k, qs, ks,
                                                       cu_seqlens_q_scale, cu_seqlens_k_scale, odt, causal), q)
        cq, ck = _cu32(cu_seqlens_q, dev), _cu32(cu_seqlens_k, dev)
        cqs, cks = _cu32(cu_seqlens_q_scale, dev), _cu32(cu_seqlens_k_scale, dev)
        assert vt.dtype == torch.float16, "V must be float16 for the FP16 P.V path"
        o = _attend_packed(qt, kt, vt, qs.t().contiguous(), ks.t().contiguous(), cq, ck, cqs, cks, max_seqlen_q, odt,
                           causal)
    return T.like(o, q)


def _forward_varlen_per_sequence(qt, kt, vt, cu_seqlens_q, cu_seqlens_k, qs, ks, cu_qs, cu_ks, odt, causal):
    """compat_tail path: every sequence is the strided view [1, n_i, H, D] of the packed tensors and goes through the
    padded kernel entry (TMA zero-fills past the view's end, like the reference's masked loads); lengths on the host."""
    cq, ck = T.as_torch(cu_seqlens_q).tolist(), T.as_torch(cu_seqlens_k).tolist()
    cqs, cks = T.as_torch(cu_qs).tolist(), T.as_torch(cu_ks).tolist()
    o = torch.empty(qt.shape, dtype=odt, device=qt.device)
    for i in range(len(cq) - 1):
        a, b, c, e = cq[i], cq[i + 1], ck[i], ck[i + 1]
        if b == a:
            continue
        if e == c:
            o[a:b].zero_()
            continue
        qs_i = qs[cqs[i]:cqs[i + 1]].t().contiguous().unsqueeze(0)
        ks_i = ks[cks[i]:cks[i + 1]].t().contiguous().unsqueeze(0)
        A._forward(qt[a:b].unsqueeze(0), kt[c:e].unsqueeze(0), vt[c:e].unsqueeze(0), qs_i, ks_i, "NHD", odt, False,
                   causal, compat_tail=True, out=o[a:b].unsqueeze(0))
    return o


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
    qt, kt, vt = T.aligned16(qt), T.aligned16(kt), T.aligned16(vt)
    backend = kwargs.get("quantization_backend", "triton")
    if backend not in Qz._MODES:
        raise ValueError(f"Unsupported quantization backend: {backend}")
    if kwargs.get("compat_tail", False):
        raise ValueError("compat_tail is only available through forward_varlen (per-sequence path)")
    with torch.cuda.device(dev):
        if dtype == torch.bfloat16:
            vt = vt.to(torch.float16)
        cq, ck = _cu32(cu_seqlens_q, dev), _cu32(cu_seqlens_k, dev)
        assert cq.numel() == ck.numel(), "cu_seqlens_q and cu_seqlens_k must describe the same batch"
        km = k_mean_varlen(kt).reshape(kt.shape[1], d_to) if smooth_k else None
        if sm_scale is None:
            sm_scale = 1.0 / head_dim_og ** 0.5
        mode = Qz._MODES[backend]
        k_c, k_s, cks, _ = _quant_packed(kt, km, ck, max_seqlen_k, 64, 8, False, 1.0, mode)
        q_c, q_s, cqs, _ = _quant_packed(qt, None, cq, max_seqlen_q, 128, 8, False, sm_scale * Qz.LOG2E, mode)
        o = _attend_packed(q_c, k_c, vt, q_s, k_s, cq, ck, cqs, cks, max_seqlen_q, dtype, bool(is_causal))
    return T.like(o[..., :head_dim_og], q)


lowbit_fa_varlen = sageattn_varlen
