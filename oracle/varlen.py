"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's variable-length (packed cu_seqlens) path.

Follows (paths relative to the reference repository):
  sageattn_varlen                    src/core.py:356-491 (glue: pad :428-438, bf16 V -> fp16 :445-446,
                                     batch-GLOBAL K mean `k.mean(dim=0)` + `k - km` :447-449, quantize :452-461,
                                     attention :462-489, slice :490)
  quant_per_block_int8_kernel        src/triton/quant_per_block_varlen.py:22-72 (Q1 arithmetic, blocks restart at every
                                     sequence, scales packed block-major [total_blocks, H] :64)
  per_block_int8 (varlen host)       src/triton/quant_per_block_varlen.py:75-142
  _attn_fwd / forward (varlen)       src/triton/attn_qk_int8_block_varlen.py:24-248 (= A1 per sequence; tail keys are
                                     zero-filled but NOT masked :72-75, i.e. the compat_tail behaviour), causal twin
                                     src/triton/attn_qk_int8_per_block_causal_varlen.py

Pinned: tests/golden/varlen_*.npz hold inputs and outputs of the reference's own varlen kernels run under the
Triton interpreter (tools/make_golden_varlen.py); tests/test_oracle_pinned.py checks this restatement against
them (codes / scales bit-exact, attention <= 4 fp16 ulps).  Built on oracle/quant.py and oracle/attention.py, one
sequence at a time.  Only tests/ may import this module.
"""
import torch

from . import attention as OA
from . import quant as OQ


def _cu_scale(cu, blk):
    out = [0]
    for a, b in zip(cu, cu[1:]):
        out.append(out[-1] + (b - a + blk - 1) // blk)
    return out


def k_mean_varlen(k):
    """core.py:448 `k.mean(dim=0, keepdim=True)` under the km contract of oracle.quant.k_mean: [1,H,D]."""
    return OQ.k_mean(k.unsqueeze(0), "NHD").reshape(1, k.shape[1], k.shape[2])


def per_block_int8_varlen(q, k, cu_q, cu_k, BLKQ=128, BLKK=64, sm_scale=None):
    """k is already smoothed (the reference subtracts before the call).  cu_*: python lists.
    -> (q_int8, q_scale [nblk,Hq], k_int8, k_scale [nblk,Hkv], cu_q_scale, cu_k_scale)."""
    d = q.shape[-1]
    if sm_scale is None:
        sm_scale = d ** -0.5
    res = []
    for x, cu, blk, sm in ((q, cu_q, BLKQ, sm_scale * 1.44269504), (k, cu_k, BLKK, 1.0)):
        cs = _cu_scale(cu, blk)
        codes = torch.zeros(x.shape, dtype=torch.int8)
        scale = torch.zeros((cs[-1], x.shape[1]), dtype=torch.float32)
        for i in range(len(cu) - 1):
            a, b = cu[i], cu[i + 1]
            if b == a:
                continue
            c, s = OQ.quant_per_block_q1(x[a:b].unsqueeze(0), blk, sm, "NHD", 8)
            codes[a:b] = c[0]
            scale[cs[i]:cs[i + 1]] = s[0].t()
        res += [codes, scale, cs]
    return res[0], res[1], res[3], res[4], res[2], res[5]


def attn_varlen(q_c, k_c, v, cu_q, cu_k, q_s, k_s, cu_qs, cu_ks, causal=False, output_dtype=torch.float16,
                pv_accum="fp16_block", compat_tail=True):
    o = torch.zeros(q_c.shape, dtype=output_dtype)
    for i in range(len(cu_q) - 1):
        a, b, c, e = cu_q[i], cu_q[i + 1], cu_k[i], cu_k[i + 1]
        if b == a or e == c:
            continue
        qs_i = q_s[cu_qs[i]:cu_qs[i + 1]].t().contiguous().unsqueeze(0)
        ks_i = k_s[cu_ks[i]:cu_ks[i + 1]].t().contiguous().unsqueeze(0)
        oi, _ = OA.attn_block_emulator(q_c[a:b].unsqueeze(0), k_c[c:e].unsqueeze(0), v[c:e].unsqueeze(0), qs_i, ks_i,
                                       "NHD", causal, output_dtype, False, pv_accum, compat_tail)
        o[a:b] = oi[0]
    return o


def lowbit_fa_varlen_api(q, k, v, cu_q, cu_k, is_causal=False, sm_scale=None, smooth_k=True, pv_accum="fp16_block",
                         compat_tail=True):
    """core.py:356-491 end to end (cu_*: python lists)."""
    dtype = q.dtype
    d_og = q.shape[-1]
    d_to = 64 if d_og <= 64 else 128
    q, k, v = (torch.nn.functional.pad(t, (0, d_to - d_og)) if d_og != d_to else t for t in (q, k, v))
    if dtype == torch.bfloat16:
        v = v.to(torch.float16)
    if smooth_k:
        k = k - k_mean_varlen(k)
    if sm_scale is None:
        sm_scale = 1.0 / d_og ** 0.5
    q_c, q_s, k_c, k_s, cqs, cks = per_block_int8_varlen(q, k, cu_q, cu_k, sm_scale=sm_scale)
    o = attn_varlen(q_c, k_c, v, cu_q, cu_k, q_s, k_s, cqs, cks, is_causal, dtype, pv_accum, compat_tail)
    return o[..., :d_og]
