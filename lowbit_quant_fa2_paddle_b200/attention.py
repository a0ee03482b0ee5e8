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


# INT4 K codes travel packed (two per byte, low nibble = even d) and are expanded to int8 in shared memory by the
# kernel (qk_mode QK_Q8K4); one-code-per-int8 K codes go through qk_mode QK_I8 (identical arithmetic).
PACKED_K4_KERNEL = True


def _out_dtype(output_dtype, default):
    if output_dtype is None:
        return default
    name = str(output_dtype).split(".")[-1]
    return {"float16": torch.float16, "bfloat16": torch.bfloat16}.get(name, default)


def _v_strides(vt, tensor_layout, pv_mode):
    """(stride_b, stride_h, stride_n) of fp16 V, or the byte strides of (b, h, d) of the e4m3 V^T that
    per_channel_fp8 returns ([B,H,D,Npad] for HND, [B,D,H,Npad] for NHD; src/quant.py:262-274)."""
    if pv_mode == N.PV_F16:
        assert vt.dtype == torch.float16, "V must be float16 for the FP16 P.V path"
        return T.bhnd(vt, tensor_layout)[4:]
    assert vt.dtype == torch.float8_e4m3fn and vt.dim() == 4 and vt.stride(3) == 1
    if tensor_layout == "HND":
        return vt.stride(0), vt.stride(1), vt.stride(2)
    return vt.stride(0), vt.stride(2), vt.stride(1)


def _common(q, k, v, q_scale, k_scale, tensor_layout, qk_mode, pv_mode, v_scale, v_mean, kbits=None):
    if tensor_layout not in ("HND", "NHD"):
        raise ValueError(f"tensor_layout {tensor_layout} not supported")
    qt, kt, vt = T.as_torch(q), T.as_torch(k), T.as_torch(v)
    qs = T.as_torch(q_scale)
    ks = T.as_torch(k_scale) if k_scale is not None else qs  # QK_F16: one scale for every score, k_scale unused
    dev = T.require_cuda(qt, kt, vt, qs, ks)
    b, hq, nq, d, qsb, qsh, qsn = T.bhnd(qt, tensor_layout)
    _, hkv, nk, dk, ksb, ksh, ksn = T.bhnd(kt, tensor_layout)
    assert dk == (d // 2 if qk_mode == N.QK_Q8K4 else d), "K codes have the wrong last dimension for this qk_mode"
    vsb, vsh, vsn = _v_strides(vt, tensor_layout, pv_mode)
    if qk_mode == N.QK_F16:
        assert qt.dtype == kt.dtype and qt.dtype in (torch.float16, torch.bfloat16), "QK_F16 takes fp16 / bf16 q and k"
    else:
        assert qt.dtype == torch.int8 and kt.dtype == torch.int8
    assert qs.dtype == torch.float32 and ks.dtype == torch.float32 and qs.is_contiguous() and ks.is_contiguous()
    vs = vm = None
    if pv_mode == N.PV_E4M3:
        assert v_scale is not None, "the FP8 P.V path needs v_scale"
        vs = T.as_torch(v_scale).contiguous()
        vm = T.as_torch(v_mean).contiguous() if v_mean is not None else None
        assert vs.dtype == torch.float32 and tuple(vs.shape) == (b, hkv, d)
    kb = None
    if qk_mode == N.QK_Q8KMIX:
        assert kbits is not None, "mixed-width K needs the kbits map returned by per_block_k_mixed"
        kb = T.as_torch(kbits)
        assert kb.dtype == torch.int32 and kb.is_contiguous() and tuple(kb.shape) == (b, hkv, (nk + 63) // 64)
    ptrs = (qt.data_ptr(), kt.data_ptr(), vt.data_ptr(), qs.data_ptr(), ks.data_ptr(),
            vs.data_ptr() if vs is not None else None, vm.data_ptr() if vm is not None else None,
            kb.data_ptr() if kb is not None else None)
    dims = (b, hq, hkv, nq, nk, d, qsb, qsh, qsn, ksb, ksh, ksn, vsb, vsh, vsn)
    return dev, qt, ptrs, dims, (vs, vm, kb)


def _forward(q, k, v, q_scale, k_scale, tensor_layout, output_dtype, return_lse, causal,
             qk_mode=N.QK_I8, pv_mode=N.PV_F16, compat_tail=False, v_scale=None, v_mean=None, kbits=None, out=None,
             narrow=False):
    dev, qt, ptrs, dims, keep = _common(q, k, v, q_scale, k_scale, tensor_layout, qk_mode, pv_mode, v_scale, v_mean, kbits)
    b, hq, hkv, nq, nk, d = dims[:6]
    if causal:
        assert nq == nk, "qo_len and kv_len must be equal for causal attention"
    odt = _out_dtype(output_dtype, torch.float16)
    if out is None:
        o = torch.empty(qt.shape, dtype=odt, device=dev)
    else:  # caller-owned output (e.g. one sequence's rows of a packed varlen tensor)
        o = out
        assert o.shape == qt.shape and o.dtype == odt and o.device == dev and o.stride(-1) == 1
    _, _, _, _, osb, osh, osn = T.bhnd(o, tensor_layout)
    lse = torch.empty((b, hq, nq), dtype=torch.float32, device=dev) if return_lse else None
    flags = (N.ATTN_CAUSAL if causal else 0) | (N.ATTN_COMPAT_TAIL if compat_tail else 0) | (N.ATTN_NARROW if narrow else 0)
    N.call("lowbit_attn_fwd", *ptrs, o.data_ptr(), lse.data_ptr() if lse is not None else None,
           *dims, osb, osh, osn, qk_mode, pv_mode, T.dtype_code(odt), flags, T.stream_ptr(dev), device=dev)
    if lse is None:
        lse = torch.empty([0], dtype=torch.float32)  # the reference returns an empty CPU tensor (:204)
    return T.like(o, q), T.like(lse, q)


class PartialState:
    """Un-normalised running attention state of a ring / sequence-parallel pass over K/V shards:
    m [B,Hq,Nq] (base-2 running maximum), l [B,Hq,Nq], o_acc [B,Hq,Nq,D] fp32, all in dequantized units."""

    def __init__(self, b, hq, nq, d, device):
        self.m = torch.empty((b, hq, nq), dtype=torch.float32, device=device)
        self.l = torch.empty((b, hq, nq), dtype=torch.float32, device=device)
        self.o_acc = torch.empty((b, hq, nq, d), dtype=torch.float32, device=device)
        self.started = False


def forward_partial(state, q, k, v, q_scale, k_scale, tensor_layout="HND", causal=False, q_offset=0, k_offset=0,
                    qk_mode=N.QK_I8, pv_mode=N.PV_F16, v_scale=None, v_mean=None, kbits=None):
    """One ring step: attend the resident Q shard to one K/V shard and merge into `state` (created on first use
    when None).  q_offset / k_offset: global token positions of the shards' first rows (causal masking)."""
    dev, qt, ptrs, dims, keep = _common(q, k, v, q_scale, k_scale, tensor_layout, qk_mode, pv_mode, v_scale, v_mean, kbits)
    b, hq, hkv, nq, nk, d = dims[:6]
    if state is None:
        state = PartialState(b, hq, nq, d, dev)
    N.call("lowbit_attn_fwd_partial", *ptrs, state.m.data_ptr(), state.l.data_ptr(), state.o_acc.data_ptr(),
           *dims, int(q_offset), int(k_offset), qk_mode, pv_mode, N.ATTN_CAUSAL if causal else 0,
           0 if state.started else 1, T.stream_ptr(dev), device=dev)
    state.started = True
    return state


def finalize(state, like_q, tensor_layout="HND", output_dtype=torch.float16, return_lse=False):
    """o = o_acc / l in the layout of `like_q` (the Q codes tensor), lse2 = log2(l) + m."""
    qt = T.as_torch(like_q)
    dev = qt.device
    b, hq, nq, d = state.o_acc.shape
    odt = _out_dtype(output_dtype, torch.float16)
    o = torch.empty(qt.shape, dtype=odt, device=dev)
    _, _, _, _, osb, osh, osn = T.bhnd(o, tensor_layout)
    lse = torch.empty((b, hq, nq), dtype=torch.float32, device=dev) if return_lse else None
    N.call("lowbit_attn_finalize", state.m.data_ptr(), state.l.data_ptr(), state.o_acc.data_ptr(), o.data_ptr(),
           lse.data_ptr() if lse is not None else None, b, hq, nq, d, osb, osh, osn, T.dtype_code(odt),
           T.stream_ptr(dev), device=dev)
    return o, lse


def forward(q, k, v, q_scale, k_scale, tensor_layout="HND", output_dtype=torch.float16, return_lse=False,
            compat_tail=False, **modes):
    """Non-causal INT8-QK / FP16-PV attention over pre-quantized codes.  `modes` (qk_mode, pv_mode, v_scale,
    v_mean) select packed INT4 K codes and the FP8 P.V path; narrow=True forces the 32-key-step kernel at
    head_dim 64 (LOWBIT_ATTN_NARROW)."""
    return _forward(q, k, v, q_scale, k_scale, tensor_layout, output_dtype, return_lse, False,
                    compat_tail=compat_tail, **modes)


def forward_causal(q, k, v, q_scale, k_scale, tensor_layout="HND", output_dtype=torch.float16, return_lse=False,
                   **modes):
    """Causal INT8-QK / FP16-PV attention over pre-quantized codes (qo_len == kv_len)."""
    return _forward(q, k, v, q_scale, k_scale, tensor_layout, output_dtype, return_lse, True, **modes)
