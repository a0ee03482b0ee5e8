"""Public operator API -- drop-in for the reference's `src/core.py` low-bit attention entry points.

  lowbit_fa_qk_int8_pv_fp16_triton   src/core.py:194-352  (alias :1102)
  lowbit_fa_qk_int4_pv_fp16_triton   src/core.py:945-1036 (alias :1105)
  lowbit_fa_q_int8_k_int4_pv_fp16    mixed entry: quant_per_block.py:391-458 + utils/paddle_package.py:321-417
  lowbit_fa_multi_precision          src/core.py:1064-1096 (select_quantization :1050-1061)
  lowbit_fa_attn (sageattn)          src/core.py:82-190   (alias :1099) -- plug-and-play dispatcher
  lowbit_fa_qk_int8_pv_fp16_cuda     src/core.py:495-731  (alias :1103) -- CUDA-API signature on the sm_100a kernel
  lowbit_fa_qk_int8_pv_fp8_cuda      src/core.py:735-941  (alias :1104) -- FP8 P.V semantics on the sm_100a kernel
  lowbit_fa_qk_int4_pv_fp8           INT4 K + FP8 P.V (BASELINE config 3)

Same names, keyword arguments, return values, assertion / ValueError behaviour.  The "_triton" suffix is kept
because callers import these names; the implementation is hand-written sm_100a CUDA (csrc/) reached through
the C ABI -- there is no Triton, no backend dispatch and no CPU fallback.  `quantization_backend` selects the
reference's two *rounding conventions* ("triton": Q1, "cuda": Q2), both executed by the same CUDA kernel;
"triton_gpu" is Q1 with the approximate fp32 division (PTX div.full.f32) that Triton emits when it JIT-compiles the
reference kernels for a GPU -- bit-identical to those kernels on a B200, whereas "triton" is the IEEE arithmetic of
the same kernels under the Triton interpreter (what the golden vectors pin).
"""
import os
from typing import Any, Optional

import torch

from . import _native as N
from . import _tensor as T
from . import attention as A
from . import quant as Qz

LOG2E = 1.44269504


def _pad_head(x, to):
    d = x.shape[-1]
    return x if d == to else torch.nn.functional.pad(x, (0, to - d))


# The INT8 / packed-INT4 K, FP16 P.V operator goes through ONE C-ABI call (csrc/op.cu: same kernels, same arguments,
# bit-identical results; LOWBIT_ONE_CALL=0 keeps the five-call path below, which serves every other format).
_ONE_CALL = os.environ.get("LOWBIT_ONE_CALL", "1") != "0"


def _lowbit_fa_one_call(q, qt, kt, vt, dev, tensor_layout, quantization_backend, is_causal, sm_scale, smooth_k, return_lse,
                        qk, compat_tail):
    dtype = qt.dtype
    if dtype == torch.bfloat16:
        vt = vt.to(torch.float16)
    b, hq, nq, d, qsb, qsh, qsn = T.bhnd(qt, tensor_layout)
    _, hkv, nk, _, ksb, ksh, ksn = T.bhnd(kt, tensor_layout)
    _, _, _, _, vsb, vsh, vsn = T.bhnd(vt, tensor_layout)
    if is_causal:
        assert nq == nk, "qo_len and kv_len must be equal for causal attention"
    if sm_scale is None:
        sm_scale = 1.0 / d ** 0.5
    kbits, packed = (8, 0) if qk == "int8" else (4, 1)
    lib = N.lib()
    o = torch.empty(qt.shape, dtype=dtype, device=dev)
    _, _, _, _, osb, osh, osn = T.bhnd(o, tensor_layout)
    lse = torch.empty((b, hq, nq), dtype=torch.float32, device=dev) if return_lse else None
    ws = torch.empty(lib.lowbit_fa_fwd_workspace_bytes(b, hq, hkv, nq, nk, d, kbits, packed), dtype=torch.uint8, device=dev)
    flags = (N.ATTN_CAUSAL if is_causal else 0) | (N.ATTN_COMPAT_TAIL if compat_tail else 0)
    N.call("lowbit_fa_fwd", qt.data_ptr(), kt.data_ptr(), vt.data_ptr(), o.data_ptr(),
           lse.data_ptr() if lse is not None else None, ws.data_ptr(), b, hq, hkv, nq, nk, d,
           1 if tensor_layout == "NHD" else 0, qsb, qsh, qsn, ksb, ksh, ksn, vsb, vsh, vsn, osb, osh, osn,
           float(sm_scale), float(sm_scale * LOG2E), kbits, packed, int(bool(smooth_k)), Qz._MODES[quantization_backend],
           T.dtype_code(dtype), T.dtype_code(dtype), flags, T.stream_ptr(dev), device=dev)
    if return_lse:
        return T.like(o, q), T.like(lse, q)
    return T.like(o, q)


def _lowbit_fa(q, k, v, tensor_layout, quantization_backend, is_causal, sm_scale, smooth_k, return_lse,
               qk, compat_tail=False, pv="fp16", smooth_v=False, kmix=None):
    qt, kt, vt = T.as_torch(q), T.as_torch(k), T.as_torch(v)
    dtype = qt.dtype
    assert dtype in [torch.float16, torch.bfloat16], \
        "Input tensors must be in dtype of torch.float16 or torch.bfloat16"
    assert qt.device == kt.device == vt.device, "All tensors must be on the same device."
    assert qt.dtype == kt.dtype == vt.dtype, "All tensors must have the same dtype."
    if tensor_layout not in ("HND", "NHD"):
        raise ValueError(f"Unknown tensor layout: {tensor_layout}")
    if quantization_backend not in ("triton", "triton_gpu", "cuda"):
        raise ValueError(f"Unsupported quantization backend: {quantization_backend}")
    head_dim_og = qt.shape[-1]
    if head_dim_og > 128:
        raise ValueError(f"Unsupported head_dim: {head_dim_og}")
    dev = T.require_cuda(qt, kt, vt)
    d_to = 64 if head_dim_og <= 64 else 128
    qt, kt, vt = _pad_head(qt, d_to), _pad_head(kt, d_to), _pad_head(vt, d_to)
    assert qt.stride(-1) == 1 and kt.stride(-1) == 1 and vt.stride(-1) == 1, "Last dim of qkv must be contiguous."
    qt, kt, vt = T.aligned16(qt), T.aligned16(kt), T.aligned16(vt)
    if (pv == "fp16" and qk in ("int8", "int4", "q8k4") and head_dim_og in (64, 128) and _ONE_CALL
            and (qk == "int8" or A.PACKED_K4_KERNEL) and os.environ.get("LOWBIT_K_FUSED", "0") != "1"):
        return _lowbit_fa_one_call(q, qt, kt, vt, dev, tensor_layout, quantization_backend, is_causal, sm_scale, smooth_k,
                                   return_lse, qk, compat_tail)
    with torch.cuda.device(dev):
        v_scale = v_mean = None
        if pv == "fp8":  # V -> e4m3 per channel, transposed (src/quant.py:210-291; core.py:882-884)
            vt, v_scale, v_mean = Qz.per_channel_fp8(vt, tensor_layout=tensor_layout, smooth_v=smooth_v)
        elif dtype == torch.bfloat16:
            vt = vt.to(torch.float16)
        if sm_scale is None:
            sm_scale = 1.0 / head_dim_og ** 0.5
        kbits = 8 if qk == "int8" else 4
        packed = (qk != "int8") and A.PACKED_K4_KERNEL
        # K mean, then both quantizers with the K smoothing fused (core.py:291-319)
        kb = None
        if qk == "mixed":  # dynamic INT8 / INT4 / INT2 per 64-row K block
            km = Qz.k_mean(kt, tensor_layout) if smooth_k else None
            k_c, k_s, kb = Qz.per_block_k_mixed(kt, km, kmix[0], kmix[1], kmix[2], tensor_layout, quantization_backend)
            q_c, q_s = Qz._quant_one(qt, None, 128, 8, False, sm_scale * LOG2E, Qz._MODES[quantization_backend],
                                     tensor_layout)
            qk_mode = N.QK_Q8KMIX
        else:
            q_c, q_s, k_c, k_s, km = Qz.smooth_and_quantize(qt, kt, smooth_k, sm_scale, tensor_layout, 8, kbits, packed,
                                                            quantization_backend)
            qk_mode = N.QK_Q8K4 if packed else N.QK_I8
        o, lse = A._forward(q_c, k_c, vt, q_s, k_s, tensor_layout, dtype, return_lse, bool(is_causal),
                            qk_mode=qk_mode, pv_mode=N.PV_E4M3 if pv == "fp8" else N.PV_F16,
                            compat_tail=compat_tail, v_scale=v_scale, v_mean=v_mean, kbits=kb)
        o = o[..., :head_dim_og]
        if return_lse:
            b, hq, nq, d, sb, sh, sn = T.bhnd(qt, tensor_layout)
            hkv = T.bhnd(kt, tensor_layout)[1]
            kmp = Qz._km_bhd(km, b, hkv, d, tensor_layout) if smooth_k else None
            N.call("lowbit_lse_fixup", lse.data_ptr(), qt.data_ptr(), kmp.data_ptr() if kmp is not None else None,
                   b, hq, hkv, nq, d, sb, sh, sn, float(sm_scale), T.dtype_code(dtype), T.stream_ptr(dev), device=dev)
            return T.like(o, q), T.like(lse, q)
    return T.like(o, q)


def sageattn_qk_int8_pv_fp16_triton(q, k, v, tensor_layout: str = "HND", quantization_backend: str = "triton",
                                    is_causal: bool = False, sm_scale: Optional[float] = None,
                                    smooth_k: bool = True, return_lse: bool = False, **kwargs: Any):
    """Per-block INT8 Q.K^T (Q blocks of 128, K blocks of 64, K mean-smoothed) + FP16 P.V.
    q: [B,Hq,Nq,D] (HND) or [B,Nq,Hq,D] (NHD), k/v likewise with Hkv heads; fp16 or bf16.
    Returns o (same shape/dtype as q) or (o, lse [B,Hq,Nq] f32 natural log) when return_lse."""
    return _lowbit_fa(q, k, v, tensor_layout, quantization_backend, is_causal, sm_scale, smooth_k, return_lse,
                      "int8", compat_tail=bool(kwargs.get("compat_tail", False)))


def sageattn_qk_int4_pv_fp16_triton(q, k, v, tensor_layout: str = "HND", quantization_backend: str = "triton",
                                    is_causal: bool = False, sm_scale: Optional[float] = None,
                                    smooth_k: bool = True, return_lse: bool = False, **kwargs: Any):
    """INT4 entry point (core.py:945-1036).  The reference quantizes Q to 8 bit and K to 4 bit (:999-1004);
    coherent semantics per SURVEY 2.3-A: Q INT8 per block, K symmetric INT4 per 64-row block (codes in
    [-7,7], packed two per byte in HBM, unpacked to int8 in shared memory so QK^T stays exact)."""
    return _lowbit_fa(q, k, v, tensor_layout, quantization_backend, is_causal, sm_scale, smooth_k, return_lse,
                      "int4", compat_tail=bool(kwargs.get("compat_tail", False)))


def lowbit_fa_q_int8_k_int4_pv_fp16(q, k, v, tensor_layout: str = "HND", quantization_backend: str = "triton",
                                    is_causal: bool = False, sm_scale: Optional[float] = None,
                                    smooth_k: bool = True, return_lse: bool = False, **kwargs: Any):
    """Mixed q_int8 / k_int4 entry point (quantizer quant_per_block.py:391-458; the reference ships the
    harness utils/paddle_package.py:321-417 but no kernel)."""
    return _lowbit_fa(q, k, v, tensor_layout, quantization_backend, is_causal, sm_scale, smooth_k, return_lse,
                      "q8k4", compat_tail=bool(kwargs.get("compat_tail", False)))


def sageattn_qk_int8_pv_fp16_cuda(q, k, v, tensor_layout: str = "HND", is_causal: bool = False,
                                  qk_quant_gran: str = "per_thread", sm_scale: Optional[float] = None,
                                  pv_accum_dtype: str = "fp32", smooth_k: bool = True, smooth_v: bool = False,
                                  return_lse: bool = False, **kwargs: Any):
    """The signature of src/core.py:495-731 (INT8 Q.K^T + FP16 P.V, "CUDA" entry point) on the sm_100a kernel.
    The kernel dequantizes with per-block scales and always accumulates P.V in fp32, so `qk_quant_gran`
    ("per_warp" | "per_thread") and `pv_accum_dtype` ("fp16" | "fp16+fp32" | "fp32") are validated like the reference
    (:598-604) and every value is served by the fp32-accumulating kernel; `smooth_v` only exists to protect fp16
    accumulators (:684-691) and is ignored, as the reference does for "fp32" (:676-679).  Quantizer rounding follows
    the reference's CUDA quantizer (Q2: RNE, reciprocal multiply, src/quant.py:21-98) unless
    quantization_backend= is passed."""
    if qk_quant_gran not in ("per_warp", "per_thread"):
        raise ValueError(f"Unsupported qk_quant_gran: {qk_quant_gran}")
    if pv_accum_dtype not in ("fp16", "fp16+fp32", "fp32"):
        raise ValueError(f"Unsupported pv_accum_dtype: {pv_accum_dtype}")
    return _lowbit_fa(q, k, v, tensor_layout, kwargs.get("quantization_backend", "cuda"), is_causal, sm_scale,
                      smooth_k, return_lse, "int8")


def sageattn(q, k, v, tensor_layout: str = "HND", is_causal: bool = False, sm_scale: Optional[float] = None,
             return_lse: bool = False, **kwargs: Any):
    """The plug-and-play entry point (src/core.py:82-190, alias lowbit_fa_attn :1099), e.g. as a replacement for
    scaled_dot_product_attention (example/sageattn_cogvideo.py:9-14): extra SDPA keywords (attn_mask, dropout_p, ...)
    are accepted and ignored like the reference's **kwargs.  The reference picks a kernel per compute capability
    (sm80 / sm86 / sm89 / sm90) and raises for anything else; this build has one target, sm_100a, served by the
    INT8-QK / FP16-PV (fp32 accumulate) kernel -- the choice the reference makes for sm80 and sm90."""
    qt = T.as_torch(q)
    if qt.device.type == "cuda":
        cap = torch.cuda.get_device_capability(qt.device)
        if cap[0] != 10:
            raise ValueError(f"Unsupported CUDA architecture: sm{cap[0]}{cap[1]}")
    return sageattn_qk_int8_pv_fp16_cuda(q, k, v, tensor_layout=tensor_layout, is_causal=is_causal, sm_scale=sm_scale,
                                         return_lse=return_lse, pv_accum_dtype="fp32")


def sageattn_qk_int8_pv_fp8_cuda(q, k, v, tensor_layout: str = "HND", is_causal: bool = False,
                                 qk_quant_gran: str = "per_thread", sm_scale: Optional[float] = None,
                                 pv_accum_dtype: str = "fp32+fp32", smooth_k: bool = True, smooth_v: bool = False,
                                 return_lse: bool = False, **kwargs: Any):
    """INT8 Q.K^T + FP8 (e4m3) P.V with fp32 accumulation -- the signature of src/core.py:735-941.  V is quantized
    per channel by per_channel_fp8 (:882-884); P~ = e4m3(exp2(s - m + offset)), the denominator sums the rounded P~
    (attn_utils.cuh:424-428,550-562), the epilogue applies v_scale (+ v_mean when smooth_v).  The sm_100a kernel
    dequantizes with per-block scales, so `qk_quant_gran` is accepted for signature compatibility and the
    quantizer is the per-block one; `pv_accum_dtype` "fp32+fp32" ignores smooth_v like the reference (:878-881)."""
    if qk_quant_gran not in ("per_warp", "per_thread"):
        raise ValueError(f"Unsupported qk_quant_gran: {qk_quant_gran}")
    if pv_accum_dtype not in ("fp32", "fp32+fp32"):
        raise ValueError(f"Unsupported pv_accum_dtype: {pv_accum_dtype}")
    if pv_accum_dtype == "fp32+fp32":
        smooth_v = False
    return _lowbit_fa(q, k, v, tensor_layout, kwargs.get("quantization_backend", "triton"), is_causal, sm_scale,
                      smooth_k, return_lse, "int8", pv="fp8", smooth_v=smooth_v)


def lowbit_fa_qk_int4_pv_fp8(q, k, v, tensor_layout: str = "HND", is_causal: bool = False,
                             sm_scale: Optional[float] = None, smooth_k: bool = True, smooth_v: bool = False,
                             return_lse: bool = False, **kwargs: Any):
    """INT4 K (packed, per 64-row block) x INT8 Q + FP8 (e4m3) P.V: the combination BASELINE config 3 names.
    The reference has the two halves (core.py:945-1036 and :735-941) but no kernel that combines them."""
    return _lowbit_fa(q, k, v, tensor_layout, kwargs.get("quantization_backend", "triton"), is_causal, sm_scale,
                      smooth_k, return_lse, "int4", pv="fp8", smooth_v=smooth_v)


def lowbit_fa_q_int8_k_dynamic(q, k, v, tensor_layout: str = "HND", is_causal: bool = False,
                               sm_scale: Optional[float] = None, smooth_k: bool = True, return_lse: bool = False,
                               kbits=None, hi: float = 0.2, lo: float = 0.05, pv: str = "fp16", **kwargs: Any):
    """Dynamic K bit allocation (BASELINE config 5): Q INT8 per 128-row block, every 64-row K block INT8,
    INT4 or INT2 by the thresholds of select_quantization (core.py:1055-1061) applied to the block scale
    max|k - km| / 127 (> lo: 8 bits, > lo/4: 4, else 2), or by an explicit `kbits` map [B,Hkv,ceil(N/64)].
    The attention kernel loads D*bits/8 bytes per K row and expands them to int8 in shared memory, so the integer
    products stay exact.  pv: "fp16" | "fp8"."""
    if pv not in ("fp16", "fp8"):
        raise ValueError(f"Unsupported pv: {pv}")
    return _lowbit_fa(q, k, v, tensor_layout, kwargs.get("quantization_backend", "triton"), is_causal, sm_scale,
                      smooth_k, return_lse, "mixed", pv=pv, kmix=(kbits, hi, lo))


def compute_scale(tensor, bits=8, symmetric=True, tensor_layout="HND"):
    """core.py:1039-1047 -- symmetric: max|x| / (2^(bits-1) - 1); asymmetric: (max - min) / (2^bits - 1); a 0-d device
    tensor either way."""
    t = T.as_torch(tensor)
    d = t.shape[-1]
    if d > 128:
        raise ValueError(f"Unsupported head_dim: {d}")
    if not symmetric:
        if d not in (64, 128):  # pad by repeating the last channel: zeros would change the extremes
            seg = t[..., -1:].expand(*t.shape[:-1], (64 if d <= 64 else 128) - d)
            t = torch.cat([t, seg], dim=-1)
        mm = Qz.min_max(T.aligned16(t.contiguous() if t.stride(-1) != 1 else t), tensor_layout)
        return (mm[0] - mm[1]) / (2 ** bits - 1)
    if d not in (64, 128):
        t = _pad_head(t, 64 if d <= 64 else 128)
    return Qz.abs_max(T.aligned16(t), tensor_layout) / (2 ** (bits - 1) - 1)


def select_quantization(q, k, v, tensor_layout="HND"):
    """core.py:1050-1061: mean of the three global scales -> "FP16" (>0.2) | "INT8" (>0.05) | "INT4".
    One device->host read, like the reference."""
    avg = (compute_scale(q, 8, True, tensor_layout) + compute_scale(k, 8, True, tensor_layout)
           + compute_scale(v, 8, True, tensor_layout)) / 3.0
    avg = float(avg)
    if avg > 0.2:
        return "FP16"
    if avg > 0.05:
        return "INT8"
    return "INT4"


_scale_cache = {}  # (device index, value) -> 1-element f32 device tensor holding sm_scale * log2(e)


def lowbit_fa_fp16(q, k, v, tensor_layout: str = "HND", is_causal: bool = False, sm_scale: Optional[float] = None,
                   return_lse: bool = False, **kwargs: Any):
    """Un-quantized attention on the same kernel: the "FP16" class of `lowbit_fa_multi_precision`, which the reference
    serves with `default_attn` (plain scaled-dot-product attention, src/core.py:46-69,1075-1076).  Q.K^T runs on
    tcgen05 kind::f16 over the fp16 / bf16 inputs as they are (fp32 scores), the rest of the pipeline (base-2 online
    softmax, fp16 P.V with fp32 accumulation) is shared with the low-bit modes; no K smoothing (softmax is invariant
    under it and nothing is quantized).  Returns o, or (o, lse [B,Hq,Nq] natural log) with return_lse."""
    qt, kt, vt = T.as_torch(q), T.as_torch(k), T.as_torch(v)
    dtype = qt.dtype
    assert dtype in [torch.float16, torch.bfloat16], \
        "Input tensors must be in dtype of torch.float16 or torch.bfloat16"
    assert qt.device == kt.device == vt.device, "All tensors must be on the same device."
    assert qt.dtype == kt.dtype == vt.dtype, "All tensors must have the same dtype."
    if tensor_layout not in ("HND", "NHD"):
        raise ValueError(f"Unknown tensor layout: {tensor_layout}")
    head_dim_og = qt.shape[-1]
    if head_dim_og > 128:
        raise ValueError(f"Unsupported head_dim: {head_dim_og}")
    dev = T.require_cuda(qt, kt, vt)
    d_to = 64 if head_dim_og <= 64 else 128
    qt, kt, vt = (T.aligned16(_pad_head(t, d_to)) for t in (qt, kt, vt))
    assert qt.stride(-1) == 1 and kt.stride(-1) == 1 and vt.stride(-1) == 1, "Last dim of qkv must be contiguous."
    if sm_scale is None:
        sm_scale = 1.0 / head_dim_og ** 0.5
    key = (dev.index, float(sm_scale))
    sc = _scale_cache.get(key)
    if sc is None:
        if len(_scale_cache) > 64:
            _scale_cache.clear()
        sc = _scale_cache[key] = torch.full((1,), float(sm_scale) * LOG2E, dtype=torch.float32, device=dev)
    if dtype == torch.bfloat16:
        vt = vt.to(torch.float16)  # P.V runs in fp16 like every other mode (core.py:307-308)
    o, lse = A._forward(qt, kt, vt, sc, None, tensor_layout, dtype, return_lse, bool(is_causal), qk_mode=N.QK_F16)
    o = o[..., :head_dim_og]
    if return_lse:
        return T.like(o, q), T.like(lse / LOG2E, q)
    return T.like(o, q)


def sageattn_multi_precision(q, k, v, tensor_layout: str = "HND", is_causal: bool = False,
                             sm_scale: Optional[float] = None, return_lse: bool = False, **kwargs: Any):
    """core.py:1064-1096: the mean of the three global scales picks the format -- "FP16" (> 0.2): un-quantized
    attention (the reference calls `default_attn`, :1075-1076; here the same kernel with fp16 Q.K^T,
    `lowbit_fa_fp16`), "INT8" (> 0.05), else "INT4"."""
    kind = select_quantization(q, k, v, tensor_layout)
    fn = {"FP16": lowbit_fa_fp16, "INT8": sageattn_qk_int8_pv_fp16_triton, "INT4": sageattn_qk_int4_pv_fp16_triton}[kind]
    return fn(q, k, v, tensor_layout=tensor_layout, is_causal=is_causal, sm_scale=sm_scale, return_lse=return_lse)


# preferred names (core.py:1099-1105)
lowbit_fa_attn = sageattn
lowbit_fa_qk_int8_pv_fp16_cuda = sageattn_qk_int8_pv_fp16_cuda
lowbit_fa_multi_precision = sageattn_multi_precision
lowbit_fa_qk_int8_pv_fp16_triton = sageattn_qk_int8_pv_fp16_triton
lowbit_fa_qk_int4_pv_fp16_triton = sageattn_qk_int4_pv_fp16_triton
lowbit_fa_qk_int8_pv_fp8_cuda = sageattn_qk_int8_pv_fp8_cuda
