"""Host-resident operands: the same operators, streamed through the GPU in (batch, head-group) chunks.

The reference asserts that q, k, v already live on the GPU (`src/core.py:269-276`).  Callers that keep activations or
an offloaded KV cache in pinned host memory would otherwise pay  H2D(q,k,v) -> quantize -> attention -> D2H(o)
strictly one after the other; every (batch, kv-head) unit of the operator is independent (K smoothing,
the quantization blocks and the attention tiles never cross a (b, h) boundary), so `lowbit_fa_host` cuts the call
into chunks of such units and overlaps the three stages on three CUDA streams:

    copy-in stream :  H2D chunk i+1            (PCIe, host -> device)
    caller's stream:  quantize + attention i   (SMs)
    copy-out stream:  D2H chunk i-1            (PCIe, device -> host; full duplex with the copy-in)

The result is bit-identical to one call on the whole tensors (tests/test_gpu_parity.py::test_host_streaming_*).
This is what `bench.py` reports as `e2e`.
"""
from typing import Any, Callable, List, Optional, Tuple

import torch

from . import _native as N
from . import _tensor as T

_streams = {}


def _side_streams(dev: torch.device):
    key = (dev.type, dev.index)
    if key not in _streams:
        _streams[key] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
    return _streams[key]


_NSLOT = 3
_stage_cache = {}


def _staging(dev: torch.device, dtype, shapes):
    """Persistent device staging buffers (q, k, v chunk + a "slot free" event) x _NSLOT, cached per chunk geometry, so
    the pipelined loop never goes to the allocator (a cudaMalloc in the loop shows up as a multi-millisecond stall)."""
    key = (dev.index, dtype, shapes)
    slots = _stage_cache.get(key)
    if slots is None:
        if len(_stage_cache) >= 4:  # a handful of geometries at most: drop the oldest
            _stage_cache.pop(next(iter(_stage_cache)))
        slots = []
        for _ in range(_NSLOT):
            bufs = [torch.empty(sh, dtype=dtype, device=dev) for sh in shapes]
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            slots.append((bufs[0], bufs[1], bufs[2], ev))
        _stage_cache[key] = slots
    return slots


def plan_chunks(B: int, Hq: int, Hkv: int, tensor_layout: str, chunks: Optional[int]) -> List[Tuple[int, int, int]]:
    """Cut the (batch, kv-head) units into chunks whose host memory is ONE contiguous range per tensor (a strided
    host slice would need a CPU-side gather before the DMA).  Returns [(b, kv_head_begin, kv_head_end), ...].
    HND `[B,H,N,D]`: any head range of one batch entry is contiguous.  NHD `[B,N,H,D]`: only whole batch entries are.
    `chunks` is a target count (default 8); q heads follow their kv head (GQA groups are never split)."""
    if tensor_layout not in ("HND", "NHD"):
        raise ValueError(f"Unknown tensor layout: {tensor_layout}")
    if Hq % Hkv != 0:
        raise ValueError("Hq must be a multiple of Hkv")
    target = 8 if chunks is None else max(1, int(chunks))
    per_b = 1
    if tensor_layout == "HND":
        per_b = max(1, min(Hkv, -(-target // B)))
        while Hkv % per_b:  # equal head groups
            per_b -= 1
    step = Hkv // per_b
    return [(b, h0, h0 + step) for b in range(B) for h0 in range(0, Hkv, step)]


def lowbit_fa_host(q, k, v, out=None, op: Optional[Callable[..., Any]] = None, tensor_layout: str = "HND",
                   chunks: Optional[int] = None, device=None, **op_kwargs: Any):
    """Run `op` (default `lowbit_fa_qk_int8_pv_fp16_triton`) on HOST tensors q, k, v (pinned memory for asynchronous
    DMA) and return the HOST tensor `out` (allocated pinned when not given), overlapping the copies with the kernels.
    Work is ordered on the caller's current stream of `device`: when this returns, everything is enqueued and the
    current stream has been made to wait for the last copy-out -- synchronize it (or an event on it) before reading
    `out` on the host.  `return_lse` is not supported on this entry point."""
    from . import core
    op = op or core.lowbit_fa_qk_int8_pv_fp16_triton
    if op_kwargs.get("return_lse"):
        raise ValueError("lowbit_fa_host does not return lse")
    qt, kt, vt = T.as_torch(q), T.as_torch(k), T.as_torch(v)
    assert qt.device.type == "cpu" and kt.device.type == "cpu" and vt.device.type == "cpu", \
        "lowbit_fa_host takes host tensors; device tensors go to the operator directly"
    assert qt.dtype == kt.dtype == vt.dtype, "All tensors must have the same dtype."
    assert qt.is_contiguous() and kt.is_contiguous() and vt.is_contiguous(), "host tensors must be contiguous"
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise N.LowbitNativeError("lowbit_fa_host needs a CUDA device: there is no CPU fallback")
    if tensor_layout == "HND":
        B, Hq, Hkv = qt.shape[0], qt.shape[1], kt.shape[1]
    elif tensor_layout == "NHD":
        B, Hq, Hkv = qt.shape[0], qt.shape[2], kt.shape[2]
    else:
        raise ValueError(f"Unknown tensor layout: {tensor_layout}")
    if out is None:
        out = torch.empty(qt.shape, dtype=qt.dtype, pin_memory=True)
    assert out.shape == qt.shape and out.dtype == qt.dtype and out.device.type == "cpu" and out.is_contiguous()
    grp = Hq // Hkv
    plan = plan_chunks(B, Hq, Hkv, tensor_layout, chunks)

    def view(t, b, h0, h1):
        return t[b:b + 1, h0:h1] if tensor_layout == "HND" else t[b:b + 1, :, h0:h1]

    with torch.cuda.device(dev):
        cur = torch.cuda.current_stream(dev)
        s_in, s_out = _side_streams(dev)
        s_in.wait_stream(cur)
        s_out.wait_stream(cur)
        b0, h0, h1 = plan[0]
        shapes = tuple(tuple(view(t, b0, a, e).shape) for t, a, e in ((qt, h0 * grp, h1 * grp), (kt, h0, h1), (vt, h0, h1)))
        slots = _staging(dev, qt.dtype, shapes)

        def stage(i):
            """H2D of chunk i into staging slot i % NSLOT, after the kernels that last read that slot."""
            b, h0, h1 = plan[i]
            dq, dk, dv, free = slots[i % _NSLOT]
            with torch.cuda.stream(s_in):
                s_in.wait_event(free)
                dk.copy_(view(kt, b, h0, h1), non_blocking=True)  # K first: its mean + codes head the chunk
                dq.copy_(view(qt, b, h0 * grp, h1 * grp), non_blocking=True)
                dv.copy_(view(vt, b, h0, h1), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_in)
            return dq, dk, dv, ev, free

        staged = stage(0)
        for i, (b, h0, h1) in enumerate(plan):
            dq, dk, dv, ev, free = staged
            if i + 1 < len(plan):
                staged = stage(i + 1)  # enqueue the next H2D before this chunk's kernels
            cur.wait_event(ev)
            o = op(dq, dk, dv, tensor_layout=tensor_layout, **op_kwargs)
            free.record(cur)  # the slot may be overwritten once these kernels are done
            with torch.cuda.stream(s_out):
                s_out.wait_event(free)
                view(out, b, h0 * grp, h1 * grp).copy_(o, non_blocking=True)
            o.record_stream(s_out)
        cur.wait_stream(s_out)
    return out
