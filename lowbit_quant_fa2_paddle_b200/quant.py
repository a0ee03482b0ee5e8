"""Quantizers of the low-bit attention path -- host side (Python over framework tensors, kernels in
csrc/quant.cu via the C ABI).  Names, arguments, return values and error behaviour mirror the reference:

  per_block_int8            src/triton/quant_per_block.py:181-248 (backend "triton", Q1)
                            src/quant.py:21-98                    (backend "cuda",   Q2)
  per_block_int4_unpack     src/triton/quant_per_block.py:251-318
  per_block_int4            src/triton/quant_per_block.py:321-388 (intent: packed INT4)
  per_block_q_int8_k_int4   src/triton/quant_per_block.py:391-458 (intent: Q INT8 + K packed INT4)
  k_mean                    src/core.py:293

Unlike the reference, `k - km` is never materialised: the subtraction is fused into the quantize kernel.
"""
import torch

from . import _native as N
from . import _tensor as T

LOG2E = 1.44269504


def k_mean(k, tensor_layout="HND"):
    """Mean of k over the sequence dimension, keepdim (core.py:293), in k's dtype.
    fp16: exact order-independent sum -> fp32 -> /N -> fp16 (SURVEY 2.3-H contract)."""
    kt = T.as_torch(k)
    dev = T.require_cuda(kt)
    b, h, n, d, sb, sh, sn = T.bhnd(kt, tensor_layout)
    km = torch.empty((b, h, d), dtype=kt.dtype, device=dev)
    ws = torch.empty(N.lib().lowbit_k_mean_workspace_bytes(b, h, n, d), dtype=torch.uint8, device=dev)
    N.call("lowbit_k_mean", kt.data_ptr(), km.data_ptr(), ws.data_ptr(), b, h, n, d, sb, sh, sn,
           T.dtype_code(kt.dtype), T.stream_ptr(dev), device=dev)
    km = km.unsqueeze(2) if tensor_layout == "HND" else km.unsqueeze(1)
    return T.like(km, k)


def _km_bhd(km, b, h, d, tensor_layout):
    """Accept km as [B,H,1,D] / [B,1,H,D] (keepdim mean) or [B,H,D]; return contiguous [B,H,D]."""
    if km is None:
        return None
    kmt = T.as_torch(km)
    if kmt.dim() == 4:
        kmt = kmt.squeeze(2 if tensor_layout == "HND" else 1)
    assert tuple(kmt.shape) == (b, h, d), f"km must have shape [{b},{h},{d}] (got {tuple(kmt.shape)})"
    return kmt.contiguous()


def _quant_one(x, km, blk, bits, pack, sm_arg, mode, tensor_layout, out=None):
    """Quantize one [B,H,N,D] tensor per block of `blk` rows. Returns (codes, scale).
    out: optional (codes, scale) torch tensors to write into (e.g. views of a ring message buffer)."""
    xt = T.as_torch(x)
    dev = T.require_cuda(xt)
    b, h, n, d, sb, sh, sn = T.bhnd(xt, tensor_layout)
    if d not in (64, 128):
        raise ValueError(f"Unsupported head_dim: {d} (the kernels take 64 or 128; core pads smaller ones)")
    kmt = _km_bhd(km, b, h, d, tensor_layout)
    if kmt is not None:
        assert kmt.dtype == xt.dtype, "km must have the same dtype as k"
    dd = d * bits // 8 if (pack and bits < 8) else d
    shape = list(xt.shape)
    shape[-1] = dd
    nblk = (n + blk - 1) // blk
    if out is None:
        codes = torch.empty(shape, dtype=torch.int8, device=dev)
        scale = torch.empty((b, h, nblk), dtype=torch.float32, device=dev)
    else:
        codes, scale = out
        assert list(codes.shape) == shape and codes.dtype == torch.int8 and codes.stride(-1) == 1
        assert tuple(scale.shape) == (b, h, nblk) and scale.dtype == torch.float32 and scale.is_contiguous()
    _, _, _, _, osb, osh, osn = T.bhnd(codes, tensor_layout)
    N.call("lowbit_quant_per_block", xt.data_ptr(), kmt.data_ptr() if kmt is not None else None,
           codes.data_ptr(), scale.data_ptr(), b, h, n, d, sb, sh, sn, osb, osh, osn,
           blk, bits, int(bool(pack)), float(sm_arg), mode, T.dtype_code(xt.dtype), T.stream_ptr(dev), device=dev)
    return T.like(codes, x), T.like(scale, x)


def k_smooth_quant_supported(k, tensor_layout="HND"):
    """True when `k_smooth_quant` takes this tensor: fp16, head_dim 64 / 128, and a (batch, head) slice of K
    (N * D * 2 bytes) fits the shared memory of one 8-CTA cluster (~1.6 MB)."""
    kt = T.as_torch(k)
    if kt.dtype != torch.float16 or kt.dim() != 4 or kt.shape[-1] not in (64, 128):
        return False
    n = kt.shape[2] if tensor_layout == "HND" else kt.shape[1]
    return bool(N.lib().lowbit_k_smooth_quant_supported(int(n), int(kt.shape[-1]), N.F16))


def k_smooth_quant(k, bits=8, pack=False, tensor_layout="HND", backend="triton"):
    """`km = k.mean(seq)`, `k - km` and its per-64-row-block codes (core.py:291-306 + quant_per_block.py:181-248) in
    ONE launch that reads K from HBM once: a thread-block cluster per (batch, head) slice keeps the slice in shared
    memory between the exact column sum and the quantizer.  -> (km keepdim, k_codes, k_scale), bit-identical to
    `k_mean` followed by the per-block quantizer with that km."""
    if tensor_layout not in ("HND", "NHD"):
        raise ValueError(f"Unknown tensor layout: {tensor_layout}")
    if backend not in _MODES:
        raise ValueError(f"Unsupported quantization backend: {backend}")
    kt = T.as_torch(k)
    dev = T.require_cuda(kt)
    b, h, n, d, sb, sh, sn = T.bhnd(kt, tensor_layout)
    km = torch.empty((b, h, d), dtype=kt.dtype, device=dev)
    shape = list(kt.shape)
    shape[-1] = d * bits // 8 if (pack and bits < 8) else d
    codes = torch.empty(shape, dtype=torch.int8, device=dev)
    scale = torch.empty((b, h, (n + 63) // 64), dtype=torch.float32, device=dev)
    _, _, _, _, osb, osh, osn = T.bhnd(codes, tensor_layout)
    N.call("lowbit_k_smooth_quant", kt.data_ptr(), km.data_ptr(), codes.data_ptr(), scale.data_ptr(), b, h, n, d,
           sb, sh, sn, osb, osh, osn, bits, int(bool(pack)), _MODES[backend], T.dtype_code(kt.dtype), T.stream_ptr(dev), device=dev)
    km = km.unsqueeze(2) if tensor_layout == "HND" else km.unsqueeze(1)
    return T.like(km, k), T.like(codes, k), T.like(scale, k)


def _smooth_k_codes(kt, smooth_k, kbits, kpack, mode_name, tensor_layout):
    """K side of the preparation step: (km | None, k_codes, k_scale).  One cluster launch when the slice fits
    and LOWBIT_K_FUSED=1 asks for it, else k_mean + the per-block quantizer (the default: see DESIGN.md 4.1)."""
    import os
    if smooth_k and os.environ.get("LOWBIT_K_FUSED", "0") == "1" and k_smooth_quant_supported(kt, tensor_layout):
        return k_smooth_quant(kt, kbits, kpack, tensor_layout, mode_name)
    km = k_mean(kt, tensor_layout) if smooth_k else None
    k_c, k_s = _quant_one(kt, km, 64, kbits, kpack, 1.0, _MODES[mode_name], tensor_layout)
    return km, k_c, k_s


def _per_block(q, k, km, BLKQ, BLKK, sm_scale, tensor_layout, qbits, kbits, kpack, backend):
    if tensor_layout not in ("HND", "NHD"):
        raise ValueError(f"Unknown tensor layout: {tensor_layout}")
    if backend not in _MODES:  # "triton_gpu": Q1 with the JIT-compiled kernels' approximate division (div.full.f32)
        raise ValueError(f"Unsupported quantization backend: {backend}")
    mode = _MODES[backend]
    head_dim = T.as_torch(q).shape[-1]
    if sm_scale is None:
        sm_scale = head_dim ** -0.5
    # K first: it was just read by k_mean and is still (partly) L2-resident
    k_c, k_s = _quant_one(k, km, BLKK, kbits, kpack, 1.0, mode, tensor_layout)
    q_c, q_s = _quant_one(q, None, BLKQ, qbits, False, sm_scale * LOG2E, mode, tensor_layout)
    return q_c, q_s, k_c, k_s


_side = {}  # device index -> side stream for the Q quantizer


def smooth_and_quantize(q, k, smooth_k, sm_scale, tensor_layout, qbits, kbits, kpack, backend, overlap=None):
    """`km = k.mean(seq)`, K codes from `k - km`, Q codes (core.py:291-319).  The Q quantizer does not depend on the K
    chain (mean -> K codes), and each of these 15-30 us HBM-bound launches leaves the chip half idle while it ramps up
    and drains, so Q runs on a side stream next to the K chain and the caller's stream waits for it before the
    attention launch (LOWBIT_QUANT_OVERLAP=0 serialises them).  -> (q_codes, q_scale, k_codes, k_scale, km)."""
    import os
    if tensor_layout not in ("HND", "NHD"):
        raise ValueError(f"Unknown tensor layout: {tensor_layout}")
    if backend not in _MODES:
        raise ValueError(f"Unsupported quantization backend: {backend}")
    mode = _MODES[backend]
    qt, kt = T.as_torch(q), T.as_torch(k)
    dev = T.require_cuda(qt, kt)
    if sm_scale is None:
        sm_scale = qt.shape[-1] ** -0.5
    if overlap is None:
        overlap = os.environ.get("LOWBIT_QUANT_OVERLAP", "1") != "0"
    if not overlap:
        km, k_c, k_s = _smooth_k_codes(kt, smooth_k, kbits, kpack, backend, tensor_layout)
        q_c, q_s = _quant_one(qt, None, 128, qbits, False, sm_scale * LOG2E, mode, tensor_layout)
        return q_c, q_s, k_c, k_s, km
    b, h, n, d, _, _, _ = T.bhnd(qt, tensor_layout)
    q_c = torch.empty(qt.shape, dtype=torch.int8, device=dev)
    q_s = torch.empty((b, h, (n + 127) // 128), dtype=torch.float32, device=dev)
    cur = torch.cuda.current_stream(dev)
    side = _side.get(dev.index)
    if side is None:
        side = _side[dev.index] = torch.cuda.Stream(dev)
    fork, join = torch.cuda.Event(), torch.cuda.Event()
    fork.record(cur)
    side.wait_event(fork)
    with torch.cuda.stream(side):
        _quant_one(qt, None, 128, qbits, False, sm_scale * LOG2E, mode, tensor_layout, out=(q_c, q_s))
        join.record(side)
    km, k_c, k_s = _smooth_k_codes(kt, smooth_k, kbits, kpack, backend, tensor_layout)
    cur.wait_event(join)
    return q_c, q_s, k_c, k_s, km


_MODES = {"triton": N.QMODE_TRITON, "triton_gpu": N.QMODE_TRITON | N.QMODE_FLAG_DIV_FULL, "cuda": N.QMODE_CUDA}

def per_block_int8(q, k, km=None, BLKQ=128, BLKK=64, sm_scale=None, tensor_layout="HND", backend="triton"):
    """-> (q_int8, q_scale [B,Hq,ceil(Nq/BLKQ)] f32, k_int8, k_scale [B,Hkv,ceil(Nk/BLKK)] f32).
    Q is scaled by sm_scale*1.44269504 before quantization; K is smoothed by km when given."""
    return _per_block(q, k, km, BLKQ, BLKK, sm_scale, tensor_layout, 8, 8, False, backend)


def per_block_int8_cuda(q, k, km=None, BLKQ=128, BLKK=64, sm_scale=None, tensor_layout="HND"):
    """src/quant.py:21-98 conventions (RNE, reciprocal multiply, amax floor 1e-7, fp32 mean subtraction)."""
    return _per_block(q, k, km, BLKQ, BLKK, sm_scale, tensor_layout, 8, 8, False, "cuda")


def per_block_int4_unpack(q, k, km=None, BLKQ=128, BLKK=64, sm_scale=None, tensor_layout="HND", backend="triton"):
    """quant_per_block.py:251-318: both Q and K to INT4 codes in [-7,7], one code per int8."""
    return _per_block(q, k, km, BLKQ, BLKK, sm_scale, tensor_layout, 4, 4, False, backend)


def per_block_int4(q, k, km=None, BLKQ=128, BLKK=64, sm_scale=None, tensor_layout="HND"):
    """Intent of quant_per_block.py:321-388: Q and K INT4, K packed two codes per byte
    along head_dim (low nibble = even d); Q codes stay unpacked for the int8 tensor-core operand."""
    return _per_block(q, k, km, BLKQ, BLKK, sm_scale, tensor_layout, 4, 4, True, "triton")


def per_block_q_int8_k_int4(q, k, km=None, BLKQ=128, BLKK=64, sm_scale=None, tensor_layout="HND", pack=True):
    """Intent of quant_per_block.py:391-458: Q INT8 per 128-row block, K INT4 per 64-row block,
    K packed two codes per byte (pack=False keeps one code per int8)."""
    return _per_block(q, k, km, BLKQ, BLKK, sm_scale, tensor_layout, 8, 4, pack, "triton")


def per_block_k_lowbit(k, km=None, BLKK=64, bits=2, tensor_layout="HND", pack=True):
    """K only, INT2/INT4/INT8 symmetric per block (dynamic bit allocation building block)."""
    return _quant_one(k, km, BLKK, bits, pack, 1.0, N.QMODE_TRITON, tensor_layout)


def per_block_k_mixed(k, km=None, kbits=None, hi=0.2, lo=0.05, tensor_layout="HND", backend="triton", out=None):
    """Dynamic K bit allocation: every 64-row block of k - km goes to INT8, INT4 or INT2.
    kbits: optional int32 [B,H,ceil(N/64)] map to impose; else the block statistic st = max|k - km| / 127
    (compute_scale, core.py:1039-1047) picks 8 bits when st > lo (both upper classes of select_quantization,
    core.py:1055-1061: there is no per-block FP16 fallback), 4 when st > lo / 4, else 2.
    -> (codes: int8 container [.., D] -- a block of width w uses the first D*w/8 bytes of its rows --,
        k_scale f32 [B,H,nblk], kbits int32 [B,H,nblk]); attention consumes it with qk_mode QK_Q8KMIX."""
    if tensor_layout not in ("HND", "NHD"):
        raise ValueError(f"Unknown tensor layout: {tensor_layout}")
    if backend not in _MODES:
        raise ValueError(f"Unsupported quantization backend: {backend}")
    kt = T.as_torch(k)
    dev = T.require_cuda(kt)
    b, h, n, d, sb, sh, sn = T.bhnd(kt, tensor_layout)
    if d not in (64, 128):
        raise ValueError(f"Unsupported head_dim: {d} (the kernels take 64 or 128; core pads smaller ones)")
    kmt = _km_bhd(km, b, h, d, tensor_layout)
    nblk = (n + 63) // 64
    if out is None:
        codes = torch.zeros(kt.shape, dtype=torch.int8, device=dev)  # unused row tails stay zero
        scale = torch.empty((b, h, nblk), dtype=torch.float32, device=dev)
        bits = torch.empty((b, h, nblk), dtype=torch.int32, device=dev)
    else:
        codes, scale, bits = out
    kin = None
    if kbits is not None:
        kin = T.as_torch(kbits).to(device=dev, dtype=torch.int32).contiguous()
        assert tuple(kin.shape) == (b, h, nblk), f"kbits must have shape [{b},{h},{nblk}]"
    _, _, _, _, osb, osh, osn = T.bhnd(codes, tensor_layout)
    thr8 = float(torch.tensor(lo, dtype=torch.float32))
    thr4 = float(torch.tensor(lo / 4, dtype=torch.float32))
    N.call("lowbit_quant_k_mixed", kt.data_ptr(), kmt.data_ptr() if kmt is not None else None,
           kin.data_ptr() if kin is not None else None, codes.data_ptr(), scale.data_ptr(), bits.data_ptr(),
           b, h, n, d, sb, sh, sn, osb, osh, osn, thr8, thr4, _MODES[backend], T.dtype_code(kt.dtype), T.stream_ptr(dev), device=dev)
    return T.like(codes, k), T.like(scale, k), T.like(bits, k)


def _per_thread(q, k, km, BLKQ, BLKK, WARPQ, WARPK, tensor_layout, bits):
    if tensor_layout not in ("HND", "NHD"):
        raise ValueError(f"Unknown tensor layout: {tensor_layout}")
    outs = []
    for x, kmx, warp_blk, blk, is_key in ((q, None, WARPQ, BLKQ, 0), (k, km, WARPK, BLKK, 1)):
        xt = T.as_torch(x)
        dev = T.require_cuda(xt)
        b, h, n, d, sb, sh, sn = T.bhnd(xt, tensor_layout)
        if d not in (64, 128):
            raise ValueError(f"Unsupported head_dim: {d}")
        kmt = _km_bhd(kmx, b, h, d, tensor_layout)
        codes = torch.empty(xt.shape, dtype=torch.int8, device=dev)
        _, _, _, _, osb, osh, osn = T.bhnd(codes, tensor_layout)
        n_scale = (n + blk - 1) // blk * (blk // warp_blk) * (4 if is_key else 8)  # quant_per_thread.py:269-278
        scale = torch.empty((b, h, n_scale), dtype=torch.float32, device=dev)
        N.call("lowbit_quant_per_thread", xt.data_ptr(), kmt.data_ptr() if kmt is not None else None,
               codes.data_ptr(), scale.data_ptr(), b, h, n, d, sb, sh, sn, osb, osh, osn, warp_blk, n_scale, is_key,
               bits, T.dtype_code(xt.dtype), T.stream_ptr(dev), device=dev)
        outs += [T.like(codes, x), T.like(scale, x)]
    return tuple(outs)


def per_thread_int8(q, k, km=None, BLKQ=128, BLKK=64, WARPQ=32, WARPK=64, sm_scale=None, tensor_layout="HND"):
    """src/triton/quant_per_thread.py:222-315: scales per mma-fragment row group -- Q [B,H,ceil(N/128)*4*8],
    K [B,Hkv,ceil(N/64)*4]; scale = amax/127 + 1e-7; sm_scale is NOT folded (the attention kernel applies it)."""
    return _per_thread(q, k, km, BLKQ, BLKK, WARPQ, WARPK, tensor_layout, 8)


def per_thread_int4(q, k, km=None, BLKQ=128, BLKK=64, WARPQ=32, WARPK=64, sm_scale=None, tensor_layout="HND"):
    """src/triton/quant_per_thread.py:317-411: as per_thread_int8 with QMAX 7, one code per int8."""
    return _per_thread(q, k, km, BLKQ, BLKK, WARPQ, WARPK, tensor_layout, 4)


def per_warp_int8(q, k, km=None, tensor_layout="HND", sm_scale=None):
    """src/quant.py:101-172: Q per 32-row warp block (scale [B,H,ceil(N/128)*4]), K per 64-row block, CUDA
    rounding conventions (Q2); sm_scale is not folded."""
    qt = T.as_torch(q)
    b, h, n, d, *_ = T.bhnd(qt, tensor_layout)
    q_c, q_s = _quant_one(q, None, 32, 8, False, 1.0, N.QMODE_CUDA, tensor_layout)
    want = (n + 127) // 128 * 4
    if q_s.shape[2] < want:  # warp blocks wholly past the data: amax floor 1e-7
        pad = torch.full((b, h, want - q_s.shape[2]), 1e-7, dtype=torch.float32, device=T.as_torch(q_s).device) / 127.0
        q_s = T.like(torch.cat([T.as_torch(q_s), pad], dim=2), q)
    k_c, k_s = _quant_one(k, km, 64, 8, False, 1.0, N.QMODE_CUDA, tensor_layout)
    return q_c, q_s, k_c, k_s


def triton_quantize_and_pack_along_last_dim(data, group_size: int, bit: int):
    """src/triton/utils/quant/new_pack.py:247-300 (KIVI): asymmetric per-group (32) quantization along the last
    dim in fp16 arithmetic, packed 8/bit codes per byte.  data [B, D, nh, T] fp16 ->
    (code int8 [B,D,nh,T*bit/8], scale fp16 [B,D,nh,T/32], mn fp16 [B,D,nh,T/32])."""
    dt = T.as_torch(data)
    assert dt.dim() == 4
    dev = T.require_cuda(dt)
    B, D, nh, Tn = dt.shape
    assert Tn % group_size == 0
    assert dt.dtype == torch.float16, "KIVI pack operates on float16 data"
    dt = dt.contiguous()
    ng = Tn // group_size
    code = torch.empty((B, D, nh, Tn * bit // 8), dtype=torch.int8, device=dev)
    scale = torch.empty((B, D, nh, ng), dtype=torch.float16, device=dev)
    mn = torch.empty((B, D, nh, ng), dtype=torch.float16, device=dev)
    N.call("lowbit_quant_pack_lastdim", dt.data_ptr(), code.data_ptr(), scale.data_ptr(), mn.data_ptr(),
           B * D * nh, Tn, group_size, bit, N.F16, T.stream_ptr(dev), device=dev)
    return T.like(code, data), T.like(scale, data), T.like(mn, data)


def per_channel_fp8(v, tensor_layout="HND", scale_max=448.0, smooth_v=True):
    """src/quant.py:210-291: V -> float8_e4m3fn per channel, transposed to [B,H,D,Npad64] (HND) / [B,D,H,Npad64]
    (NHD), tokens permuted inside 16-groups; returns (v_fp8, v_scale [B,H,D] f32, vm [B,H,D] f32 | None)."""
    vt = T.as_torch(v)
    dev = T.require_cuda(vt)
    b, h, n, d, sb, sh, sn = T.bhnd(vt, tensor_layout)
    if d not in (64, 128):
        raise ValueError(f"Unsupported head_dim: {d}")
    npad = (n + 63) // 64 * 64
    if tensor_layout == "HND":
        v8 = torch.empty((b, h, d, npad), dtype=torch.float8_e4m3fn, device=dev)
        osb, osh, osd = v8.stride(0), v8.stride(1), v8.stride(2)
    else:
        v8 = torch.empty((b, d, h, npad), dtype=torch.float8_e4m3fn, device=dev)
        osb, osd, osh = v8.stride(0), v8.stride(1), v8.stride(2)
    v_scale = torch.empty((b, h, d), dtype=torch.float32, device=dev)
    vm = torch.empty((b, h, d), dtype=torch.float32, device=dev) if smooth_v else None
    ws = torch.empty(N.lib().lowbit_v_fp8_workspace_bytes(b, h, n, d), dtype=torch.uint8, device=dev)
    N.call("lowbit_v_fp8_per_channel", vt.data_ptr(), v8.data_ptr(), v_scale.data_ptr(),
           vm.data_ptr() if vm is not None else None, ws.data_ptr(), b, h, n, d, sb, sh, sn, osb, osh, osd,
           float(scale_max), T.dtype_code(vt.dtype), T.stream_ptr(dev), device=dev)
    return T.like(v8, v), T.like(v_scale, v), T.like(vm, v)


def sub_mean(v, tensor_layout="HND"):
    """src/quant.py:175-207: (v_smoothed fp16 = v - mean_n(v), vm [B,H,D] in v's dtype).  The mean is this package's
    exact K-mean kernel (fp16: exact sum, one rounding; see k_mean); the subtraction follows SubMeanKernel
    (fused.cu:243-248): taken in the input dtype, then converted to fp16."""
    vt = T.as_torch(v)
    dev = T.require_cuda(vt)
    if tensor_layout not in ("HND", "NHD"):
        raise ValueError(f"Unknown tensor layout: {tensor_layout}")
    b, h, n, d, sb, sh, sn = T.bhnd(vt, tensor_layout)
    vm = _km_bhd(k_mean(vt, tensor_layout), b, h, d, tensor_layout)
    return sub_mean_given(vt, vm, tensor_layout), T.like(vm, v)


def sub_mean_given(v, vm, tensor_layout="HND"):
    """The SubMeanKernel step alone for a mean the caller hands over ([B,H,D], v's dtype): fp16(v - vm)."""
    vt, vmt = T.as_torch(v), T.as_torch(vm).contiguous()
    dev = T.require_cuda(vt, vmt)
    b, h, n, d, sb, sh, sn = T.bhnd(vt, tensor_layout)
    assert tuple(vmt.shape) == (b, h, d) and vmt.dtype == vt.dtype, "vm must be [B,H,D] in v's dtype"
    if d % 8 != 0:
        raise ValueError(f"Unsupported head_dim: {d}")
    out = torch.empty(vt.shape, dtype=torch.float16, device=dev)
    _, _, _, _, osb, osh, osn = T.bhnd(out, tensor_layout)
    N.call("lowbit_sub_mean", vt.data_ptr(), vmt.data_ptr(), out.data_ptr(), b, h, n, d, sb, sh, sn, osb, osh, osn,
           T.dtype_code(vt.dtype), T.stream_ptr(dev), device=dev)
    return T.like(out, v)


def abs_max(x, tensor_layout="HND"):
    """Global max|x| as a 0-d device tensor (compute_scale numerator, core.py:1039-1047)."""
    xt = T.as_torch(x)
    dev = T.require_cuda(xt)
    b, h, n, d, sb, sh, sn = T.bhnd(xt, tensor_layout)
    out = torch.empty((), dtype=torch.float32, device=dev)
    N.call("lowbit_abs_max", xt.data_ptr(), out.data_ptr(), b, h, n, d, sb, sh, sn, T.dtype_code(xt.dtype),
           T.stream_ptr(dev), device=dev)
    return out


def min_max(x, tensor_layout="HND"):
    """Global (max, min) of x as a 2-element fp32 device tensor (asymmetric compute_scale, core.py:1043-1045)."""
    xt = T.as_torch(x)
    dev = T.require_cuda(xt)
    b, h, n, d, sb, sh, sn = T.bhnd(xt, tensor_layout)
    out = torch.empty((2,), dtype=torch.float32, device=dev)
    N.call("lowbit_min_max", xt.data_ptr(), out.data_ptr(), b, h, n, d, sb, sh, sn, T.dtype_code(xt.dtype),
           T.stream_ptr(dev), device=dev)
    return out
