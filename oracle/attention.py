"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference attention kernels and API glue.

Restates (paths relative to /root/reference):
  A1  src/triton/attn_qk_int8_per_block.py:24-167          (non-causal _attn_fwd/_attn_fwd_inner)
  A2  src/triton/attn_qk_int8_per_block_causal.py:24-79,216-334   (causal _attn_fwd_base)
  A3  csrc/qattn/attn_utils.cuh:30,424-428,550-562 + qk_int_sv_f8_cuda.cu:557-600   (FP8 PV; spec only)
  E1  src/core.py:269-352 (API glue: pad, K smoothing, sm_scale, LSE fix-up)
  cpu_baseline: src/core.py:46-69 manual_scaled_dot_product_attention with the K transpose fixed
                 over dequantized Q1 codes.

Pinned by tests/test_oracle_pinned.py against tests/golden/attn_*.npz, which hold outputs of the
reference's own kernels run under the Triton interpreter (tools/make_golden.py).  The reference
accumulates P.V in fp16 inside a 64-key block (`tl.dot(..., out_dtype=fp16)`); `pv_accum="fp16_block"`
models that, `pv_accum="fp32"` is what the sm_100a kernel does.  Agreement with the goldens is by
tolerance (exp2 / fp16 dot ordering are Triton-interpreter internals), stated in the tests.
"""
import math

import torch

from . import quant as Q

LOG2E = Q.LOG2E
FP8_OFFSET = 8.807  # attn_utils.cuh:30


def _hnd(x, layout):
    return Q._to_hnd(x, layout)


def sdpa_fp32(q, k, v, tensor_layout="HND", is_causal=False, sm_scale=None, return_lse=False):
    """Plain fp32 softmax(q k^T * sm_scale) v with GQA; the accuracy yardstick (not the reference)."""
    qh, kh, vh = (_hnd(t, tensor_layout).float() for t in (q, k, v))
    b, hq, nq, d = qh.shape
    hkv = kh.shape[1]
    g = hq // hkv
    if g > 1:
        kh = kh.repeat_interleave(g, dim=1)
        vh = vh.repeat_interleave(g, dim=1)
    if sm_scale is None:
        sm_scale = 1.0 / math.sqrt(d)
    s = torch.matmul(qh, kh.transpose(2, 3)) * sm_scale
    if is_causal:
        nk = kh.shape[2]
        mask = torch.ones(nq, nk, dtype=torch.bool).tril()
        s = s.masked_fill(~mask, float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    o = torch.matmul(torch.softmax(s, dim=-1), vh)
    if tensor_layout == "NHD":
        o = o.permute(0, 2, 1, 3)
    return (o, lse) if return_lse else o


def attn_block_emulator(q_codes, k_codes, v, q_scale, k_scale, tensor_layout="HND", causal=False,
                        output_dtype=None, return_lse=False, pv_accum="fp16_block", compat_tail=True,
                        pv_mode="f16", v_scale=None, v_mean=None, BLOCK_M=128, BLOCK_N=64,
                        q_scale_rows=None):
    """Block-faithful emulator of A1/A2 (and A3 when pv_mode == 'e4m3').

    q_codes/k_codes int8 (unpacked codes), v fp16 [.., Nk, D] (pv_mode f16) or already-dequantizable
    fp8 values as float (pv_mode e4m3: pass v as the float tensor of e4m3 values, v_scale [B,H,D]).
    compat_tail=True reproduces the reference's unmasked tail (phantom keys with score 0,
    attn_qk_int8_per_block.py:48-49 -- SURVEY 2.3-E); False masks keys >= Nk with -inf.
    Returns (o in q_codes' layout, lse2 [B,Hq,Nq] base-2 or None)."""
    qc = _hnd(q_codes, tensor_layout).to(torch.int32)
    kc = _hnd(k_codes, tensor_layout).to(torch.int32)
    vh = _hnd(v, tensor_layout)
    b, hq, nq, d = qc.shape
    _, hkv, nk, _ = kc.shape
    g = hq // hkv
    output_dtype = output_dtype or torch.float16
    nmb = (nq + BLOCK_M - 1) // BLOCK_M
    nkb = (nk + BLOCK_N - 1) // BLOCK_N
    padq, padk = nmb * BLOCK_M - nq, nkb * BLOCK_N - nk
    qc = torch.nn.functional.pad(qc, (0, 0, 0, padq))
    kc = torch.nn.functional.pad(kc, (0, 0, 0, padk))
    vp = torch.nn.functional.pad(vh.float(), (0, 0, 0, padk))
    if g > 1:
        kc = kc.repeat_interleave(g, dim=1)
        vp = vp.repeat_interleave(g, dim=1)
        k_scale = k_scale.repeat_interleave(g, dim=1)
        if v_scale is not None:
            v_scale = v_scale.repeat_interleave(g, dim=1)
        if v_mean is not None:
            v_mean = v_mean.repeat_interleave(g, dim=1)
    # per-row q scale: [B,H,nmb] -> [B,H,Nq_pad,1]
    qs_row = q_scale.float().repeat_interleave(BLOCK_M, dim=2)[..., None]
    m = torch.full((b, hq, nmb * BLOCK_M, 1), float("-inf"))
    l = torch.ones_like(m)
    acc = torch.zeros(b, hq, nmb * BLOCK_M, d)
    rows = torch.arange(nmb * BLOCK_M).view(1, 1, -1, 1)
    row_blk = rows // BLOCK_M
    qf = qc.double()
    for j in range(nkb):
        c0 = j * BLOCK_N
        cols = torch.arange(c0, c0 + BLOCK_N).view(1, 1, 1, -1)
        if causal:
            # stage 1: key blocks strictly left of the row tile; stage 2: the 128-wide diagonal band
            active = (c0 < (row_blk + 1) * BLOCK_M)
            if not bool(active.any()):
                continue
        S = torch.matmul(qf, kc[:, :, c0:c0 + BLOCK_N].double().transpose(2, 3))  # exact integers
        qk = (S.float() * qs_row) * k_scale[:, :, j].float()[..., None, None]
        if causal:
            in_band = (c0 >= row_blk * BLOCK_M)
            qk = torch.where(in_band & (rows < cols), qk + torch.tensor(-1.0e6), qk)
        if not compat_tail:
            qk = torch.where(cols >= nk, torch.tensor(float("-inf")), qk)
        m_new = torch.maximum(m, qk.amax(dim=-1, keepdim=True))
        if pv_mode == "e4m3":
            p = torch.exp2(qk - m_new + torch.tensor(FP8_OFFSET))
            p = p.clamp(max=448.0).to(torch.float8_e4m3fn).float()
        else:
            p = torch.exp2(qk - m_new)
        alpha = torch.exp2(m - m_new)
        l_new = l * alpha + p.sum(dim=-1, keepdim=True)
        acc_new = acc * alpha
        vblk = vp[:, :, c0:c0 + BLOCK_N]
        if pv_mode == "e4m3":
            pv = torch.matmul(p, vblk)
        elif pv_accum == "fp16_block":
            pv = torch.matmul(p.half().float(), vblk).half().float()
        else:
            pv = torch.matmul(p.half().float(), vblk)
        acc_new = acc_new + pv
        if causal:
            m = torch.where(active, m_new, m)
            l = torch.where(active, l_new, l)
            acc = torch.where(active, acc_new, acc)
        else:
            m, l, acc = m_new, l_new, acc_new
    o = acc * (1.0 / l)
    if pv_mode == "e4m3":
        o = o * v_scale.float()[:, :, None, :]
        if v_mean is not None:
            o = o + v_mean.float()[:, :, None, :]
    o = o[:, :, :nq].to(output_dtype)
    lse2 = (torch.log2(l) + m)[:, :, :nq, 0]
    if pv_mode == "e4m3":
        lse2 = lse2 - FP8_OFFSET
    out = torch.empty(q_codes.shape, dtype=output_dtype)
    _hnd(out, tensor_layout).copy_(o)
    return out, (lse2 if return_lse else None)


def _pad_head(x, d_to):
    d = x.shape[-1]
    return x if d == d_to else torch.nn.functional.pad(x, (0, d_to - d))


def v8_to_natural(v8, n, tensor_layout="HND"):
    """Undo the transpose + within-16 token permutation of per_channel_fp8 (fused.cu:290-292): returns the e4m3
    values as float in the caller's layout, [B,H,N,D] (HND) or [B,N,H,D] (NHD)."""
    x = v8.float()
    if tensor_layout == "NHD":  # [B,D,H,Npad] -> [B,H,D,Npad]
        x = x.permute(0, 2, 1, 3)
    npad = x.shape[-1]
    inv = Q.token_perm(npad)          # dest position -> source token
    nat = torch.empty_like(x)
    nat[..., inv] = x                 # source token -> value
    nat = nat[..., :n].permute(0, 1, 3, 2)  # [B,H,N,D]
    return nat.contiguous() if tensor_layout == "HND" else nat.permute(0, 2, 1, 3).contiguous()


def lowbit_fa_api(q, k, v, tensor_layout="HND", is_causal=False, sm_scale=None, smooth_k=True,
                  return_lse=False, qk="int8", pv_accum="fp16_block", compat_tail=True, km=None,
                  pv="fp16", smooth_v=False, kmix=None):
    """E1/E2/E3 glue restated from src/core.py:269-352: head-dim pad (:277-287), km (:291-306; contract
    SURVEY 2.3-H via oracle.quant.k_mean), bf16 V -> fp16 (:307-308), sm_scale from the un-padded
    head dim (:309-310), Q1 quantize (:311-314), attention (:321-342), slice (:343), LSE fix-up (:344-350).
    qk: 'int8' (E1) | 'int4' / 'q8k4' (E2/E3 per SURVEY 2.3-A: Q INT8 per-block, K INT4 per-block)."""
    dtype = q.dtype
    assert dtype in (torch.float16, torch.bfloat16)
    d_og = q.shape[-1]
    if d_og > 128:
        raise ValueError(f"Unsupported head_dim: {d_og}")
    d_to = 64 if d_og <= 64 else 128
    q, k, v = (_pad_head(t, d_to) for t in (q, k, v))
    lse_corr = None
    if smooth_k:
        if km is None:
            km = Q.k_mean(k, tensor_layout)
        if return_lse:
            qh, kmh = _hnd(q, tensor_layout), _hnd(km, tensor_layout)
            g = qh.shape[1] // kmh.shape[1]
            kmh = kmh.repeat_interleave(g, dim=1) if g > 1 else kmh
            lse_corr = torch.matmul(qh.float(), kmh.float().transpose(2, 3)).squeeze(-1).to(dtype).float()  # matmul in the input dtype (:296-304)
    else:
        km = None
    v_scale = v_mean = None
    if pv == "fp8":  # core.py:882-884 -> A3 (attn_utils.cuh:424-428,550-562)
        n = v.shape[2] if tensor_layout == "HND" else v.shape[1]
        v8, v_scale, v_mean = Q.per_channel_fp8(v, tensor_layout, smooth_v=smooth_v)
        v = v8_to_natural(v8, n, tensor_layout)
    elif dtype == torch.bfloat16:
        v = v.to(torch.float16)
    if sm_scale is None:
        sm_scale = 1.0 / d_og ** 0.5
    if qk == "mixed":  # dynamic INT8 / INT4 / INT2 per K block (oracle.quant.quant_k_mixed; `kmix` = (kbits, hi, lo))
        qi, qs = Q.quant_per_block_q1(q, 128, sm_scale * LOG2E, tensor_layout, 8)
        kb, hi, lo = kmix if kmix is not None else (None, 0.2, 0.05)
        ki, ks, _ = Q.quant_k_mixed(k, km, kb, 64, tensor_layout, hi, lo)
    else:
        kbits = 8 if qk == "int8" else 4
        qi, qs, ki, ks = Q.per_block_int8_q1(q, k, km, sm_scale=sm_scale, tensor_layout=tensor_layout, kbits=kbits)
    o, lse2 = attn_block_emulator(qi, ki, v, qs, ks, tensor_layout, is_causal, dtype, return_lse,
                                  pv_accum=pv_accum, compat_tail=compat_tail,
                                  pv_mode="e4m3" if pv == "fp8" else "f16", v_scale=v_scale, v_mean=v_mean)
    o = o[..., :d_og]
    if return_lse:
        lse = lse2 / LOG2E
        if smooth_k:
            lse = lse + lse_corr * sm_scale
        return o, lse
    return o


def cpu_quantize_and_attend(q, k, v, tensor_layout="HND", is_causal=False, sm_scale=None, smooth_k=True, kbits=8):
    """The CPU baseline the bench times ("port"): the reference's pure-Paddle quantize-and-attend math
    on torch-CPU -- Q1 per-block INT8 quantize via tensor ops (kbits=4: K codes in [-7,7], the q_int8 / k_int4
    quantizer quant_per_block.py:391-458), dequantize, then the corrected manual_scaled_dot_product_attention
    (src/core.py:46-69, K transpose fixed per SURVEY 2.3-I) in fp32."""
    d = q.shape[-1]
    if sm_scale is None:
        sm_scale = 1.0 / d ** 0.5
    km = Q.k_mean(k, tensor_layout) if smooth_k else None
    qi, qs, ki, ks = Q.per_block_int8_q1(q, k, km, sm_scale=sm_scale, tensor_layout=tensor_layout, kbits=kbits)
    qh = _hnd(qi, tensor_layout).float() * qs.repeat_interleave(128, dim=2)[:, :, :_hnd(qi, tensor_layout).shape[2], None]
    kh = _hnd(ki, tensor_layout).float() * ks.repeat_interleave(64, dim=2)[:, :, :_hnd(ki, tensor_layout).shape[2], None]
    vh = _hnd(v, tensor_layout).float()
    g = qh.shape[1] // kh.shape[1]
    if g > 1:
        kh, vh = kh.repeat_interleave(g, dim=1), vh.repeat_interleave(g, dim=1)
    scores = torch.matmul(qh, kh.transpose(2, 3))  # already carries sm_scale*log2e
    if is_causal:
        n = scores.shape[-1]
        mask = torch.tril(torch.ones(n, n))
        scores = scores + (1 - mask) * -1e9
    w = torch.softmax(scores * (1.0 / LOG2E), dim=-1)
    o = torch.matmul(w, vh).to(q.dtype)
    return o if tensor_layout == "HND" else o.permute(0, 2, 1, 3)


# ----------------------------------------------------------------------------------------------
# ring / sequence-parallel steps (our own contract, include/lowbit_fa.h lowbit_attn_fwd_partial):
# flash-style merge of one K/V shard into an un-normalised running state, in exact-maximum fp32 math
# ----------------------------------------------------------------------------------------------
def attn_partial(state, q_codes, k_codes, v, q_scale, k_scale, tensor_layout="HND", causal=False,
                 q_offset=0, k_offset=0, v_scale=None):
    """state: None or dict(m [B,Hq,Nq], l [B,Hq,Nq], acc [B,Hq,Nq,D]).  q_codes/k_codes: int8 unpacked codes;
    v: fp16 values, or e4m3 values as float in natural [.., N, D] order with v_scale [B,Hkv,D] (the rounded
    P~ of A3 is modelled by rounding p * 2^8.807 to e4m3).  Key c is visible to row r iff
    k_offset + c <= q_offset + r."""
    qc = _hnd(q_codes, tensor_layout).double()
    kc = _hnd(k_codes, tensor_layout).double()
    vh = _hnd(v, tensor_layout).float()
    b, hq, nq, d = qc.shape
    hkv, nk = kc.shape[1], kc.shape[2]
    g = hq // hkv
    if g > 1:
        kc, vh = kc.repeat_interleave(g, dim=1), vh.repeat_interleave(g, dim=1)
        k_scale = k_scale.repeat_interleave(g, dim=1)
        v_scale = v_scale.repeat_interleave(g, dim=1) if v_scale is not None else None
    qs_row = q_scale.float().repeat_interleave(128, dim=2)[:, :, :nq, None]
    ks_col = k_scale.float().repeat_interleave(64, dim=2)[:, :, None, :nk]
    qk = (torch.matmul(qc, kc.transpose(2, 3)).float() * qs_row) * ks_col
    if causal:
        rows = torch.arange(nq).view(1, 1, -1, 1) + q_offset
        cols = torch.arange(nk).view(1, 1, 1, -1) + k_offset
        qk = torch.where(cols <= rows, qk, torch.tensor(float("-inf")))
    if state is None:
        state = dict(m=torch.full((b, hq, nq), float("-inf")), l=torch.zeros(b, hq, nq), acc=torch.zeros(b, hq, nq, d))
    m_cur = qk.amax(dim=-1)
    m_new = torch.maximum(state["m"], m_cur)
    safe = torch.where(torch.isinf(m_new), torch.zeros_like(m_new), m_new)
    p = torch.exp2(qk - safe[..., None])
    if v_scale is not None:
        p = (p * 2.0 ** FP8_OFFSET).clamp(max=448.0).to(torch.float8_e4m3fn).float() * 2.0 ** -FP8_OFFSET
        vh = vh * v_scale.float()[:, :, None, :]
    else:
        p = p.half().float()
    alpha = torch.exp2(state["m"] - safe)
    alpha = torch.where(torch.isinf(state["m"]), torch.zeros_like(alpha), alpha)
    l_new = state["l"] * alpha + (torch.exp2(qk - safe[..., None]).sum(-1) if v_scale is None else p.sum(-1))
    acc = state["acc"] * alpha[..., None] + torch.matmul(p, vh)
    return dict(m=m_new, l=l_new, acc=acc)


def attn_finalize(state, like_q, tensor_layout="HND", output_dtype=torch.float16):
    o = (state["acc"] / state["l"][..., None]).to(output_dtype)
    out = torch.empty(like_q.shape, dtype=output_dtype)
    _hnd(out, tensor_layout).copy_(o)
    return out, torch.log2(state["l"]) + state["m"]
