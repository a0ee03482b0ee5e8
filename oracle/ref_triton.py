"""TEST INFRASTRUCTURE ONLY -- loader for the reference's own @triton.jit kernels.

Runs the *unmodified* reference kernels from /root/reference under the Triton
CPU interpreter (TRITON_INTERPRET=1) with a 4-attribute `paddle` stub, exactly as
SURVEY.md Appendix B describes.  It exists so that golden vectors can be produced
from the reference itself (tools/make_golden.py) and so the numpy/torch
restatements in oracle/ can be pinned against it.

/root/reference does not exist on the GPU box: nothing under tests/ -m gpu,
bench.py or __graft_entry__.smoke() imports this module.  tools/ref_on_b200.py (a measurement
aid, not a gate) points LOWBIT_REFERENCE_ROOT at oracle/_ref/reference (staged, git-ignored,
by oracle/stage_ref.sh) with TRITON_INTERPRET=0 to JIT the same unmodified kernels on the B200;
outputs are allocated on the inputs' device for that purpose.  Nothing in the product
package imports anything under oracle/.

Reference host wrappers that are re-stated (not imported) here because they call
paddle.empty/.strides:
  src/triton/quant_per_block.py:181-248   (per_block_int8)
  src/triton/quant_per_block.py:251-318   (per_block_int4_unpack)
  src/triton/attn_qk_int8_per_block.py:169-238          (forward, non-causal)
  src/triton/attn_qk_int8_per_block_causal.py:216-334   (_attn_fwd_base, causal)
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("LOWBIT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", "triton"))


_mods = {}


def _load(rel):
    if rel in _mods:
        return _mods[rel]
    os.environ.setdefault("TRITON_INTERPRET", "1")
    import torch
    if "paddle" not in sys.modules:
        stub = types.ModuleType("paddle")
        stub.int8, stub.float16, stub.Tensor = torch.int8, torch.float16, torch.Tensor
        stub.__lowbit_stub__ = True
        sys.modules["paddle"] = stub
    path = os.path.join(REF_ROOT, rel)
    spec = importlib.util.spec_from_file_location("_ref_" + rel.replace("/", "_")[:-3], path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _mods[rel] = mod
    return mod


def _strides3(t, layout):
    s = t.stride()
    return (s[0], s[1], s[2]) if layout == "HND" else (s[0], s[2], s[1])


def _dims(t, layout):
    if layout == "HND":
        b, h, n, d = t.shape
    else:
        b, n, h, d = t.shape
    return b, h, n, d


def quant_per_block(x, blk, sm_scale_arg, layout="HND", bits=8):
    """Launch the reference quant kernel (int8 or int4-unpack) on one tensor.
    sm_scale_arg is the value passed as the kernel's `sm_scale` argument."""
    import torch
    qb = _load("src/triton/quant_per_block.py")
    kern = qb.quant_per_block_int8_kernel if bits == 8 else qb.quant_per_block_int4_unpack_kernel
    b, h, n, d = _dims(x, layout)
    out = torch.empty(x.shape, dtype=torch.int8, device=x.device)
    nblk = (n + blk - 1) // blk
    scale = torch.empty((b, h, nblk), dtype=torch.float32, device=x.device)
    kern[(nblk, h, b)](x, out, scale, n, *_strides3(x, layout), *_strides3(out, layout),
                       scale.stride(0), scale.stride(1), sm_scale=sm_scale_arg, C=d, BLK=blk)
    return out, scale


def per_block_int8(q, k, km=None, BLKQ=128, BLKK=64, sm_scale=None, tensor_layout="HND", kbits=8):
    """Re-statement of the host wrapper quant_per_block.py:181-248 around the real kernels."""
    if km is not None:
        k = k - km
    d = q.shape[-1]
    if sm_scale is None:
        sm_scale = d ** -0.5
    qi, qs = quant_per_block(q, BLKQ, sm_scale * 1.44269504, tensor_layout, 8)
    ki, ks = quant_per_block(k, BLKK, 1.0, tensor_layout, kbits)
    return qi, qs, ki, ks


def attn_forward(qi, ki, v, qs, ks, tensor_layout="HND", causal=False, output_dtype=None, return_lse=False):
    """Launch the reference attention kernel (non-causal _attn_fwd, causal _attn_fwd_base)."""
    import torch
    if causal:
        mod = _load("src/triton/attn_qk_int8_per_block_causal.py")
        kern, stage = mod._attn_fwd_base, 3
    else:
        mod = _load("src/triton/attn_qk_int8_per_block.py")
        kern, stage = mod._attn_fwd, 1
    b, hq, nq, d = _dims(qi, tensor_layout)
    _, hkv, nk, _ = _dims(ki, tensor_layout)
    output_dtype = output_dtype or v.dtype
    o = torch.empty(qi.shape, dtype=output_dtype, device=qi.device)
    lse = torch.empty((b, hq, nq), dtype=torch.float32, device=qi.device)
    kern[((nq + 127) // 128, hq, b)](
        qi, ki, v, qs, ks, o, lse,
        *_strides3(qi, tensor_layout), *_strides3(ki, tensor_layout),
        *_strides3(v, tensor_layout), *_strides3(o, tensor_layout),
        nq, nk, H=hq, num_kv_groups=hq // hkv, BLOCK_M=128, BLOCK_N=64, HEAD_DIM=d,
        STAGE=stage, RETURN_LSE=return_lse)
    return o, (lse if return_lse else None)


def kivi_pack(data, group_size, bit):
    """The reference's triton_quantize_and_pack_along_last_dim (new_pack.py:247-300) cannot be
    imported as-is (paddle tensor methods); only its two kernels are loadable."""
    return _load("src/triton/utils/quant/new_pack.py")


def per_thread(q, k, km=None, BLKQ=128, BLKK=64, WARPQ=32, WARPK=64, tensor_layout="HND", bits=8):
    """Host re-statement of quant_per_thread.py:222-315 / :317-411 around the real kernels."""
    import torch
    qt = _load("src/triton/quant_per_thread.py")
    if km is not None:
        k = k - km
    kq = qt.quant_query_per_thread_int8_kernel if bits == 8 else qt.quant_query_per_thread_int4_kernel
    kk = qt.quant_key_per_thread_int8_kernel if bits == 8 else qt.quant_key_per_thread_int4_kernel
    b, hq, nq, d = _dims(q, tensor_layout)
    _, hk, nk, _ = _dims(k, tensor_layout)
    qi = torch.zeros(q.shape, dtype=torch.int8)
    ki = torch.zeros(k.shape, dtype=torch.int8)
    gq = (nq + BLKQ - 1) // BLKQ * (BLKQ // WARPQ) * 8
    gk = (nk + BLKK - 1) // BLKK * (BLKK // WARPK) * 4
    qs = torch.empty((b, hq, gq), dtype=torch.float32)
    ks = torch.empty((b, hk, gk), dtype=torch.float32)
    kq[(gq, hq, b)](q, qi, qs, nq, *_strides3(q, tensor_layout), *_strides3(qi, tensor_layout),
                    qs.stride(0), qs.stride(1), C=d, BLK=WARPQ)
    kk[(gk, hk, b)](k, ki, ks, nk, *_strides3(k, tensor_layout), *_strides3(ki, tensor_layout),
                    ks.stride(0), ks.stride(1), C=d, BLK=WARPK)
    return qi, qs, ki, ks


def kivi_quantize_and_pack(data, group_size, bit):
    """new_pack.py:247-300 with the two real Triton kernels (_minmax :222-244, _pack :198-219) and the
    Paddle element-wise glue (:273-276) re-stated in torch (fp16 ops, round = half away from zero)."""
    import numpy as np
    import torch
    import triton
    npk = _load("src/triton/utils/quant/new_pack.py")
    B, D, nh, T = data.shape
    ng = T // group_size
    x = data.reshape(B * nh * D, ng, group_size).contiguous()
    mx = torch.empty((B * nh * D, ng), dtype=data.dtype)
    mn = torch.empty((B * nh * D, ng), dtype=data.dtype)
    BS = 128
    npk._minmax_along_last_dim[(triton.cdiv(x.shape[0] * x.shape[1], BS),)](
        x, mn, mx, x.numel(), x.shape[0], ng, group_size, BLOCK_SIZE_N=BS, num_warps=8)
    scale = (mx - mn) / (2 ** bit - 1)
    y = x - mn.unsqueeze(-1)
    y = y / scale.unsqueeze(-1)
    y = y.clamp(0, 2 ** bit - 1).float()
    y = torch.where(y >= 0, torch.floor(y + 0.5), torch.ceil(y - 0.5)).to(torch.int32)
    y = y.view(-1, T).contiguous()
    per = 8 // bit
    code = torch.zeros((int(np.prod(data.shape[:-1])), T // per), dtype=torch.int8)
    npk._pack_along_last_dim[(triton.cdiv(y.shape[0], BS), y.shape[1] // per)](
        bit, y, code, y.shape[0], y.shape[1], per, BLOCK_SIZE_N=BS, num_warps=8)
    return code.view(B, D, nh, -1), scale.reshape(B, D, nh, ng), mn.reshape(B, D, nh, ng)


# ---------------------------------------------------------------------------------------------- varlen (packed) path
def varlen_per_block_int8(q, k, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, BLKQ=128, BLKK=64,
                          sm_scale=None):
    """Host re-statement of quant_per_block_varlen.py:75-142 around the real varlen quantizer kernel (:22-72).
    q [Tq,Hq,D], k [Tk,Hkv,D] (k already smoothed by the caller, core.py:447-449)."""
    import torch
    m = _load("src/triton/quant_per_block_varlen.py")
    kern = m.quant_per_block_int8_kernel
    hq, hkv, d = q.shape[1], k.shape[1], q.shape[-1]
    b = cu_seqlens_q.shape[0] - 1
    if sm_scale is None:
        sm_scale = d ** -0.5
    outs = []
    for x, cu, blk, mx, h, sm in ((q, cu_seqlens_q, BLKQ, max_seqlen_q, hq, sm_scale * 1.44269504),
                                  (k, cu_seqlens_k, BLKK, max_seqlen_k, hkv, 1.0)):
        lens = cu[1:] - cu[:-1]
        cs = torch.nn.functional.pad(torch.cumsum((lens + blk - 1) // blk, 0), (1, 0)).to(cu.dtype)
        codes = torch.empty(x.shape, dtype=torch.int8, device=x.device)
        scale = torch.empty((int(cs[-1]), h), dtype=torch.float32, device=x.device)
        kern[((mx + blk - 1) // blk, h, b)](x, codes, scale, cu, cs, x.stride(1), x.stride(0), codes.stride(1),
                                            codes.stride(0), sm_scale=sm, H=h, C=d, BLK=blk)
        outs += [codes, scale, cs]
    return outs[0], outs[1], outs[3], outs[4], outs[2], outs[5]


def varlen_attn_forward(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, q_scale, k_scale, cu_q_scale, cu_k_scale,
                        causal=False, output_dtype=None):
    """Host re-statement of the varlen `forward` wrappers (attn_qk_int8_block_varlen.py:197-248, causal twin
    attn_qk_int8_per_block_causal_varlen.py:207-) around the real kernels."""
    import torch
    if causal:
        m = _load("src/triton/attn_qk_int8_per_block_causal_varlen.py")
        stage = 3
    else:
        m = _load("src/triton/attn_qk_int8_block_varlen.py")
        stage = 1
    o = torch.empty(q.shape, dtype=output_dtype or v.dtype, device=q.device)
    b = cu_seqlens_q.shape[0] - 1
    _, hq, d = q.shape
    _, hkv, _ = k.shape
    m._attn_fwd[((max_seqlen_q + 127) // 128, hq, b)](
        q, k, v, cu_seqlens_q, cu_seqlens_k, q_scale, k_scale, cu_q_scale, cu_k_scale, o,
        q.stride(1), q.stride(0), k.stride(1), k.stride(0), v.stride(1), v.stride(0), o.stride(1), o.stride(0),
        hq, hq // hkv, BLOCK_M=128, BLOCK_N=64, HEAD_DIM=d, STAGE=stage)
    return o
