"""TEST INFRASTRUCTURE ONLY -- CPU restatement (IEEE fp32, torch-CPU/numpy) of the reference quantizers.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this package; the product package `lowbit_quant_fa2_paddle_b200` never does (it raises if the CUDA
library is missing).

Every function cites the reference file:line it restates (paths relative to /root/reference).
All arithmetic is IEEE binary32 with one rounding per operation (no FMA contraction, no FTZ, true
division), which is what the Triton CPU interpreter executes; `tests/test_oracle_pinned.py` checks
these restatements bit-for-bit against golden vectors produced by the reference's own kernels
(tools/make_golden.py -> tests/golden/*.npz).

Pinned / unpinned status:
  Q1  per_block_int8  (Triton)         pinned   -- golden vectors from the reference kernel
  Q4  per_block_int4_unpack (Triton)   pinned   -- golden vectors from the reference kernel
  Q3  per_thread_int8/int4 (Triton)    pinned   -- golden vectors from the reference kernel
  Q5  KIVI min/max + pack (Triton)     pinned for the two kernels; the Paddle element-wise glue between
                                       them is restated from source ("parity unpinned" for that glue);
                                       kivi_unpack_and_dequant is bit-identical to the reference's
                                       unpack_and_dequant_{k,v}cache run unmodified (tests/golden/kvcache_*.npz)
  Q2  per_block_int8 (CUDA, fused.cu)  pinned (round 2) -- golden vectors from the reference's own fused.cu, compiled
  Q6  per_channel_fp8 (CUDA, fused.cu) unmodified and run on a B200 (oracle/build_ref_fused.py, tools/make_golden_fused.py,
      + sub_mean                       tests/golden/fused_*.npz, tests/test_fused_golden.py); only the fp32 block-reduction
                                       order of the smooth_v mean is left to a tolerance (2e-6 relative)
  INT2 / packed INT4 / kbits map       parity unpinned: no coherent reference kernel exists
"""
import numpy as np
import torch

LOG2E = 1.44269504  # the literal the reference uses (quant_per_block.py:226, core.py:347)

QMAX = {8: 127.0, 4: 7.0, 2: 1.0}


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
def _to_hnd(x, layout):
    """Return a [B,H,N,D] view of x for either layout (no copy)."""
    if layout == "HND":
        return x
    if layout == "NHD":
        return x.permute(0, 2, 1, 3)
    raise ValueError(f"Unknown tensor layout: {layout}")


def _f32(v):
    return torch.tensor(float(v), dtype=torch.float32)


def _blocks(x_hnd, blk):
    """[B,H,N,D] -> zero padded [B,H,nblk,blk,D] (rows >= N contribute 0, as the masked loads do)."""
    b, h, n, d = x_hnd.shape
    nblk = (n + blk - 1) // blk
    pad = nblk * blk - n
    if pad:
        x_hnd = torch.nn.functional.pad(x_hnd, (0, 0, 0, pad))
    return x_hnd.reshape(b, h, nblk, blk, d), nblk


# ----------------------------------------------------------------------------------------------
# S1: K mean (contract of SURVEY.md 2.3-H; reference call site src/core.py:293 `k.mean(dim=seq_dim)`)
# ----------------------------------------------------------------------------------------------
def k_mean(k, layout="HND"):
    """km[b,h,d] in k's dtype.  fp16: exact (order independent) sum, rounded once to fp32, divided by
    N in fp32, rounded to fp16.  bf16: sum in fp64 over fixed 64-row blocks, then fp64 across blocks in
    block order, rounded to fp32, / N, rounded to bf16.  Returned with the keepdim shape of core.py:293."""
    kh = _to_hnd(k, layout)
    b, h, n, d = kh.shape
    if k.dtype == torch.float16:
        # every finite fp16 is an integer multiple of 2^-24
        fixed = (kh.double() * float(2 ** 24)).to(torch.int64)
        s = fixed.sum(dim=2)  # exact
        s32 = s.to(torch.float32) * _f32(2.0 ** -24)  # int64->f32 RN, exact power-of-two scaling
    else:
        blk = 64
        xb, nblk = _blocks(kh.double(), blk)
        part = torch.zeros(b, h, nblk, d, dtype=torch.float64)
        for r in range(blk):  # fixed row order inside a block
            part = part + xb[:, :, :, r, :]
        s = torch.zeros(b, h, d, dtype=torch.float64)
        for j in range(nblk):  # fixed block order
            s = s + part[:, :, j, :]
        s32 = s.to(torch.float32)
    km = (s32 / _f32(n)).to(k.dtype)
    km = km.unsqueeze(2)  # [B,H,1,D]
    return km if layout == "HND" else km.permute(0, 2, 1, 3)


# ----------------------------------------------------------------------------------------------
# Q1 / Q4 / (INT2): Triton per-block symmetric quantizer
#   src/triton/quant_per_block.py:132-178 (int8), :22-71 (int4 unpack); hosts :181-248, :251-318
# ----------------------------------------------------------------------------------------------
def quant_per_block_q1(x, blk, sm_scale_arg, layout="HND", bits=8):
    """x = f32(in) * f32(sm_scale_arg); scale = max|x| / QMAX; y = x/scale; y += +-0.5; trunc.
    Returns (codes int8 in x's layout, scale f32 [B,H,nblk]).  All-zero block: scale 0, codes 0
    (the reference produces 0/0=NaN there; NaN->int8 is 0 on NVIDIA hardware)."""
    xh = _to_hnd(x, layout)
    b, h, n, d = xh.shape
    xf = xh.float() * _f32(sm_scale_arg)
    xb, nblk = _blocks(xf, blk)
    amax = xb.abs().amax(dim=(3, 4))
    scale = amax / _f32(QMAX[bits])
    y = xb / scale[..., None, None]
    y = y + torch.where(y >= 0, _f32(0.5), _f32(-0.5))
    y = torch.where(torch.isnan(y), torch.zeros_like(y), y)
    codes = y.to(torch.int32).to(torch.int8)  # trunc toward zero
    codes = codes.reshape(b, h, nblk * blk, d)[:, :, :n]
    out = torch.empty(x.shape, dtype=torch.int8)
    _to_hnd(out, layout).copy_(codes)
    return out, scale.contiguous()


def per_block_int8_q1(q, k, km=None, BLKQ=128, BLKK=64, sm_scale=None, tensor_layout="HND", kbits=8, qbits=8):
    """Host wrapper quant_per_block.py:181-248: k <- k - km in the input dtype (:186-187),
    Q scaled by sm_scale*1.44269504 (:226), K by 1.0 (:244)."""
    if km is not None:
        k = k - km  # fp16/bf16 subtraction, one rounding
    d = q.shape[-1]
    if sm_scale is None:
        sm_scale = d ** -0.5
    qi, qs = quant_per_block_q1(q, BLKQ, sm_scale * LOG2E, tensor_layout, qbits)
    ki, ks = quant_per_block_q1(k, BLKK, 1.0, tensor_layout, kbits)
    return qi, qs, ki, ks


# ----------------------------------------------------------------------------------------------
# Q2: CUDA per-block INT8   csrc/fused/fused.cu:64-198, csrc/numeric_conversion.cuh:137-142,
#     src/quant.py:70-98   (IEEE form; the reference builds with --use_fast_math)
# ----------------------------------------------------------------------------------------------
def quant_per_block_q2(x, blk, layout="HND", km=None, sm_scale_arg=None):
    """x = f32(in) [- f32(km)] [* f32(sm_scale_arg)]; amax = max(1e-7, max|x|); scale = amax/127;
    code = sat_s8(rint_half_even(x * (127/amax))).  km: [B,H,D] (squeezed, quant.py:92)."""
    xh = _to_hnd(x, layout)
    b, h, n, d = xh.shape
    xf = xh.float()
    if km is not None:
        xf = xf - km.reshape(b, h, 1, d).float()
    if sm_scale_arg is not None:
        xf = xf * _f32(sm_scale_arg)
    xb, nblk = _blocks(xf, blk)
    amax = torch.maximum(xb.abs().amax(dim=(3, 4)), _f32(1e-7))
    scale = amax / _f32(127.0)
    r = _f32(127.0) / amax
    y = torch.round(xb * r[..., None, None])  # torch.round = half to even
    codes = y.clamp(-128, 127).to(torch.int32).to(torch.int8)
    codes = codes.reshape(b, h, nblk * blk, d)[:, :, :n]
    out = torch.empty(x.shape, dtype=torch.int8)
    _to_hnd(out, layout).copy_(codes)
    return out, scale.contiguous()


def per_block_int8_q2(q, k, km=None, BLKQ=128, BLKK=64, sm_scale=None, tensor_layout="HND"):
    """src/quant.py:21-98."""
    d = q.shape[-1]
    if sm_scale is None:
        sm_scale = d ** -0.5
    sm = sm_scale * LOG2E
    qi, qs = quant_per_block_q2(q, BLKQ, tensor_layout, None, sm)
    kmm = None
    if km is not None:
        kmm = km.squeeze(1) if tensor_layout == "NHD" else km.squeeze(2)
    ki, ks = quant_per_block_q2(k, BLKK, tensor_layout, kmm, None)
    return qi, qs, ki, ks


def per_warp_int8_q2(q, k, km=None, tensor_layout="HND", sm_scale=None):
    """src/quant.py:101-172: same kernel with BLK=32 for Q (scale [B,H,ceil(N/128)*4]) and 64 for K.
    NB: the reference's per_warp path does NOT fold sm_scale (quant.py:158)."""
    qi, qs = quant_per_block_q2(q, 32, tensor_layout, None, None)
    b, h, n, d = _to_hnd(q, tensor_layout).shape
    want = (n + 127) // 128 * 4
    if qs.shape[2] < want:  # blocks that lie wholly beyond N: amax floor
        fill = torch.full((b, h, want - qs.shape[2]), 1e-7, dtype=torch.float32) / _f32(127.0)
        qs = torch.cat([qs, fill], dim=2)
    kmm = None
    if km is not None:
        kmm = km.squeeze(1) if tensor_layout == "NHD" else km.squeeze(2)
    ki, ks = quant_per_block_q2(k, 64, tensor_layout, kmm, None)
    return qi, qs, ki, ks


# ----------------------------------------------------------------------------------------------
# Q3: per-thread groups   src/triton/quant_per_thread.py:22-219 (kernels), :222-411 (hosts)
# ----------------------------------------------------------------------------------------------
def _per_thread(x, layout, bits, kind, blk):
    """kind 'q': in each `blk`(=WARPQ 32)-row block, group t in [0,8) = rows {8i+t};
    kind 'k': in each `blk`(=WARPK 64)-row block, group t in [0,4) = rows {8i+2t, 8i+2t+1}.
    scale = max|x|/QMAX + 1e-7 (:62,:114); rounding as Q1; sm_scale not folded."""
    xh = _to_hnd(x, layout)
    b, h, n, d = xh.shape
    xb, nblk = _blocks(xh.float(), blk)  # [B,H,nblk,blk,D]
    if kind == "q":
        g = xb.reshape(b, h, nblk, blk // 8, 8, d).permute(0, 1, 2, 4, 3, 5)  # [.., t(8), i, D]
        ngrp = 8
    else:
        g = xb.reshape(b, h, nblk, blk // 8, 4, 2, d).permute(0, 1, 2, 4, 3, 5, 6)  # [.., t(4), i, 2, D]
        g = g.reshape(b, h, nblk, 4, (blk // 8) * 2, d)
        ngrp = 4
    amax = g.abs().amax(dim=(4, 5))
    scale = amax / _f32(QMAX[bits]) + _f32(1e-7)
    y = g / scale[..., None, None]
    y = y + torch.where(y >= 0, _f32(0.5), _f32(-0.5))
    c = y.to(torch.int32).to(torch.int8)
    if kind == "q":
        c = c.permute(0, 1, 2, 4, 3, 5).reshape(b, h, nblk * blk, d)
    else:
        c = c.reshape(b, h, nblk, 4, blk // 8, 2, d).permute(0, 1, 2, 4, 3, 5, 6).reshape(b, h, nblk * blk, d)
    out = torch.empty(x.shape, dtype=torch.int8)
    _to_hnd(out, layout).copy_(c[:, :, :n])
    return out, scale.reshape(b, h, nblk * ngrp).contiguous()


def per_thread(q, k, km=None, BLKQ=128, BLKK=64, WARPQ=32, WARPK=64, sm_scale=None, tensor_layout="HND", bits=8):
    """quant_per_thread.py:222-315 (int8) / :317-411 (int4). Scale shapes :269-278."""
    if km is not None:
        k = k - km
    qi, qs = _per_thread(q, tensor_layout, bits, "q", WARPQ)
    ki, ks = _per_thread(k, tensor_layout, bits, "k", WARPK)
    b, h, n, d = _to_hnd(q, tensor_layout).shape
    _, hk, nk, _ = _to_hnd(k, tensor_layout).shape
    wq = (n + BLKQ - 1) // BLKQ * (BLKQ // WARPQ) * 8
    wk = (nk + BLKK - 1) // BLKK * (BLKK // WARPK) * 4
    # groups that lie wholly beyond N are still launched by the reference grid (:279,:297): amax 0 -> 1e-7
    if qs.shape[2] < wq:
        qs = torch.cat([qs, torch.full((b, h, wq - qs.shape[2]), 1e-7, dtype=torch.float32)], dim=2)
    if ks.shape[2] < wk:
        ks = torch.cat([ks, torch.full((b, hk, wk - ks.shape[2]), 1e-7, dtype=torch.float32)], dim=2)
    return qi, qs, ki, ks


# ----------------------------------------------------------------------------------------------
# packing of symmetric low-bit codes (our stated semantics, SURVEY.md 2.3-A/B/F -- parity unpinned)
# ----------------------------------------------------------------------------------------------
def pack_codes(codes, bits):
    """Pack signed codes along the last dim, little-endian inside a byte: element i of a byte at bits
    [i*bits,(i+1)*bits) as a two's-complement `bits`-wide field.  bits=4: low nibble = even d."""
    if bits == 8:
        return codes
    per = 8 // bits
    c = codes.to(torch.int32) & ((1 << bits) - 1)
    c = c.reshape(*codes.shape[:-1], codes.shape[-1] // per, per)
    out = torch.zeros(c.shape[:-1], dtype=torch.int32)
    for i in range(per):
        out |= c[..., i] << (i * bits)
    return out.to(torch.uint8).view(torch.int8)


def unpack_codes(packed, bits):
    if bits == 8:
        return packed
    per = 8 // bits
    p = packed.view(torch.uint8).to(torch.int32)
    outs = []
    for i in range(per):
        f = (p >> (i * bits)) & ((1 << bits) - 1)
        f = torch.where(f >= (1 << (bits - 1)), f - (1 << bits), f)
        outs.append(f)
    return torch.stack(outs, dim=-1).reshape(*packed.shape[:-1], packed.shape[-1] * per).to(torch.int8)


# ----------------------------------------------------------------------------------------------
# dynamic K bit-width map (our stated semantics for SURVEY 2.3-F; thresholds are core.py:1055-1061's)
# ----------------------------------------------------------------------------------------------
def k_bits_map(k_sub, blk=64, layout="HND", hi=0.2, lo=0.05):
    """kbits[b,h,j] in {8,4,2}: statistic = block max|k| / 127 (compute_scale, core.py:1039-1047);
    > hi or > lo -> 8, > lo/4 -> 4 else 2.  (FP16 fallback does not exist per block, so the top two
    classes of select_quantization both map to INT8.)"""
    kh = _to_hnd(k_sub, layout)
    xb, _ = _blocks(kh.float(), blk)
    st = xb.abs().amax(dim=(3, 4)) / _f32(127.0)
    bits = torch.full(st.shape, 2, dtype=torch.int32)
    bits[st > lo / 4] = 4
    bits[st > lo] = 8
    return bits


def quant_k_mixed(k, km=None, kbits=None, blk=64, layout="HND", hi=0.2, lo=0.05):
    """Dynamic K bit allocation (our stated semantics; parity unpinned): k <- k - km in the input dtype
    (quant_per_block.py:186-187); per 64-row block bits = kbits[b,h,j] when given else k_bits_map(k - km); codes by the
    Q1 arithmetic with QMAX[bits] (quant_per_block.py:173-176 / :60-63).
    Returns (codes int8 in k's layout, one code per element, scale f32 [B,H,nblk], kbits int32 [B,H,nblk])."""
    if km is not None:
        k = k - km
    bits = k_bits_map(k, blk, layout, hi, lo) if kbits is None else kbits.to(torch.int32)
    codes = torch.zeros(k.shape, dtype=torch.int8)
    ch = _to_hnd(codes, layout)
    scale = None
    for w in (8, 4, 2):
        c_w, s_w = quant_per_block_q1(k, blk, 1.0, layout, w)
        if scale is None:
            scale = torch.zeros_like(s_w)
        sel = bits == w                                       # [B,H,nblk]
        scale = torch.where(sel, s_w, scale)
        rows = sel.repeat_interleave(blk, dim=2)[:, :, :ch.shape[2]]
        ch.copy_(torch.where(rows[..., None], _to_hnd(c_w, layout), ch))
    return codes, scale.contiguous(), bits.contiguous()


_P4 = [0, 2, 4, 6, 1, 3, 5, 7]


def pack_mixed(codes, kbits, blk=64, layout="HND"):
    """The mixed-width container the CUDA quantizer writes (include/lowbit_fa.h, lowbit_quant_k_mixed): D bytes per
    row; a block of width w uses the first D*w/8 bytes of its rows; byte order (p = position in an 8-group g):
      8 bit: byte 8g+p = code(8g + P4[p]);   4 bit: byte 4g+i = code(8g+2i) | code(8g+2i+1) << 4;
      2 bit: byte 16c + 8(g%2) + p, bits [2(g//2), +2) = code(64c + 8g + P4[p])   (c = 64-code chunk, g = group in it).
    Unused row tails are zero.  Returns int8 in `codes`' layout."""
    ch = _to_hnd(codes, layout).to(torch.int32)
    b, h, n, d = ch.shape
    out = torch.zeros(b, h, n, d, dtype=torch.int32)
    bits_rows = kbits.repeat_interleave(blk, dim=2)[:, :, :n]
    perm = torch.tensor([8 * g + p for g in range(d // 8) for p in _P4])
    c8 = ch[..., perm]                                                       # 8-bit rows
    c4 = (ch[..., 0::2] & 15) | ((ch[..., 1::2] & 15) << 4)                   # [.., d/2]
    c2 = torch.zeros(b, h, n, d // 4, dtype=torch.int32)
    for c in range(d // 64):
        for g in range(8):
            for p in range(8):
                byte = 16 * c + 8 * (g % 2) + p
                c2[..., byte] |= (ch[..., 64 * c + 8 * g + _P4[p]] & 3) << (2 * (g // 2))
    out = torch.where((bits_rows == 8)[..., None], c8, out)
    pad4 = torch.nn.functional.pad(c4, (0, d - d // 2))
    pad2 = torch.nn.functional.pad(c2, (0, d - d // 4))
    out = torch.where((bits_rows == 4)[..., None], pad4, out)
    out = torch.where((bits_rows == 2)[..., None], pad2, out)
    res = torch.empty(codes.shape, dtype=torch.int8)
    _to_hnd(res, layout).copy_(out.to(torch.uint8).view(torch.int8))
    return res


def unpack_mixed(container, kbits, blk=64, layout="HND"):
    """Inverse of pack_mixed: one int8 code per element (what the attention kernel's in-smem expansion yields, before
    its common head-dim permutation and power-of-two factor)."""
    ch = _to_hnd(container, layout).view(torch.uint8).to(torch.int32)
    b, h, n, d = ch.shape
    bits_rows = kbits.repeat_interleave(blk, dim=2)[:, :, :n]

    def sext(f, w):
        return torch.where(f >= (1 << (w - 1)), f - (1 << w), f)

    o8 = torch.zeros(b, h, n, d, dtype=torch.int32)
    o4 = torch.zeros_like(o8)
    o2 = torch.zeros_like(o8)
    for g in range(d // 8):
        for p in range(8):
            o8[..., 8 * g + _P4[p]] = sext(ch[..., 8 * g + p], 8)
    o4[..., 0::2] = sext(ch[..., : d // 2] & 15, 4)
    o4[..., 1::2] = sext((ch[..., : d // 2] >> 4) & 15, 4)
    for c in range(d // 64):
        for g in range(8):
            for p in range(8):
                o2[..., 64 * c + 8 * g + _P4[p]] = sext((ch[..., 16 * c + 8 * (g % 2) + p] >> (2 * (g // 2))) & 3, 2)
    out = torch.where((bits_rows == 8)[..., None], o8, torch.where((bits_rows == 4)[..., None], o4, o2))
    res = torch.empty(container.shape, dtype=torch.int8)
    _to_hnd(res, layout).copy_(out.to(torch.int8))
    return res


# ----------------------------------------------------------------------------------------------
# Q5: KIVI asymmetric group quant + pack   src/triton/utils/quant/new_pack.py:198-300
# ----------------------------------------------------------------------------------------------
def kivi_quantize_and_pack(data, group_size, bit):
    """All math in the input dtype (fp16): mn,mx per group (:222-244); scale=(mx-mn)/(2^b-1) (:273);
    code = round_half_away(clip((x-mn)/scale, 0, 2^b-1)) (:274-276); pack 8/b codes per int8,
    element i at bits [i*b,(i+1)*b) (:198-219).  Returns (code [B,D,nh,T*b/8] int8, scale, mn [B,D,nh,T/g])."""
    assert data.dim() == 4
    B, D, nh, T = data.shape
    assert T % group_size == 0
    ng = T // group_size
    x = data.reshape(B * nh * D, ng, group_size)
    mx = x.amax(dim=-1)
    mn = x.amin(dim=-1)
    qm = 2 ** bit - 1
    scale = (mx - mn) / qm  # input dtype
    y = x - mn.unsqueeze(-1)
    y = y / scale.unsqueeze(-1)
    y = y.clamp(0, qm)
    yf = torch.nan_to_num(y.float(), nan=0.0)  # constant group: 0/0 in the reference; defined here as code 0
    yr = torch.where(yf >= 0, torch.floor(yf + 0.5), torch.ceil(yf - 0.5))  # Paddle round = half away
    c = yr.to(torch.int32).reshape(-1, T)
    per = 8 // bit
    c = c.reshape(c.shape[0], T // per, per)
    packed = torch.zeros(c.shape[:2], dtype=torch.int32)
    for i in range(per):
        packed |= c[..., i] << (i * bit)
    code = packed.to(torch.uint8).view(torch.int8)
    return code.reshape(B, D, nh, -1), scale.reshape(B, D, nh, ng), mn.reshape(B, D, nh, ng)


def kivi_unpack_and_dequant(code, scale, mn, group_size, bit):
    """new_pack.py:68-144 (unpack_and_dequant_*): x ~= code*scale + mn."""
    per = 8 // bit
    p = code.view(torch.uint8).to(torch.int32)
    fields = [((p >> (i * bit)) & (2 ** bit - 1)) for i in range(per)]
    c = torch.stack(fields, dim=-1).reshape(*code.shape[:-1], code.shape[-1] * per)
    c = c.reshape(*c.shape[:-1], c.shape[-1] // group_size, group_size).to(scale.dtype)
    x = c * scale.unsqueeze(-1) + mn.unsqueeze(-1)
    return x.reshape(*code.shape[:-1], -1)


# ----------------------------------------------------------------------------------------------
# Q6: V -> FP8 e4m3 per channel   src/quant.py:210-291, csrc/fused/fused.cu:263-428
# ----------------------------------------------------------------------------------------------
_PERM16 = [0, 1, 8, 9, 2, 3, 10, 11, 4, 5, 12, 13, 6, 7, 14, 15]  # fused.cu:290-292


def token_perm(npad):
    """dest position -> source token for the within-16 permutation of TransposePadPermuteKernel:
    source row r of a 16-group lands at (r/8)*2 + ((r/2)%4)*4 + (r%2)."""
    src = np.arange(npad)
    r = src % 16
    dst = (src // 16) * 16 + (r // 8) * 2 + ((r // 2) % 4) * 4 + (r % 2)
    inv = np.empty(npad, dtype=np.int64)
    inv[dst] = src
    return torch.from_numpy(inv)


def per_channel_fp8(v, tensor_layout="HND", scale_max=448.0, smooth_v=True):
    """Returns (v_fp8 [B,H,D,Npad64] float8_e4m3fn, v_scale [B,H,D] f32, vm [B,H,D] f32 | None).
    Vt padded with zeros to Npad = ceil(N/64)*64 and token-permuted inside 16-groups (fused.cu:263-314);
    statistics over ceil(N/16)*16 tokens (zero padded), mean divides by that padded count (:382);
    scale = amax/scale_max; v8 = e4m3_rn_satfinite((v [- vm]) * (scale_max/amax))."""
    vh = _to_hnd(v, tensor_layout)
    b, h, n, d = vh.shape
    npad = (n + 63) // 64 * 64
    n16 = (n + 15) // 16 * 16
    vt = torch.zeros(b, h, d, npad, dtype=v.dtype)
    vt[..., :n] = vh.permute(0, 1, 3, 2)
    vt = vt[..., token_perm(npad)]
    x = vt[..., :n16].float()
    mx = x.amax(dim=-1)
    mnv = x.amin(dim=-1)
    if smooth_v:
        # fixed-order contract for the sum: fp32 accumulation is order dependent in the reference
        # (blockReduceSum); we define it as the fp64 sum rounded to fp32 (held to the reference's fused.cu within
        # 2e-6 relative, tests/test_fused_golden.py).
        vm = (x.double().sum(dim=-1)).float() / _f32(n16)
        amax = torch.maximum((mx - vm).abs(), (mnv - vm).abs())
        xs = vt.float() - vm[..., None]
    else:
        vm = None
        amax = torch.maximum(mx.abs(), mnv.abs())
        xs = vt.float()
    v_scale = amax / _f32(scale_max)
    r = torch.where(amax > 0, _f32(scale_max) / amax, torch.zeros_like(amax))  # all-zero channel: codes 0 (reference: 0/0)
    y = (xs * r[..., None]).clamp(-448.0, 448.0)
    v8 = y.to(torch.float8_e4m3fn)
    if tensor_layout == "NHD":
        v8 = v8.permute(0, 2, 1, 3).contiguous()  # the reference allocates [B,D,H,Npad] for NHD (quant.py:269-274)
    return v8, v_scale, vm


def sub_mean(v, tensor_layout="HND"):
    """src/quant.py:175-207 -> SubMeanKernel fused.cu:201-261: vm = mean over tokens (input dtype),
    v_smoothed = fp16(v - vm).  Returns (v_smoothed fp16, vm [B,H,D] in v's dtype)."""
    seq = 1 if tensor_layout == "NHD" else 2
    vm = v.mean(dim=seq)
    return sub_mean_given(v, vm, tensor_layout)


def sub_mean_given(v, vm, tensor_layout="HND"):
    """SubMeanKernel (fused.cu:201-261) for a mean the caller hands over: the difference is taken IN THE INPUT DTYPE
    (__hsub2 on half2 / bfloat162, :243: one rounding of the exact difference) and then converted to fp16 (:245-248).
    Pinned by tests/golden/fused_*.npz (the reference kernel itself)."""
    seq = 1 if tensor_layout == "NHD" else 2
    return (v.float() - vm.unsqueeze(seq).float()).to(v.dtype).to(torch.float16), vm
