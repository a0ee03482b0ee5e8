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
import threading
from typing import Any, Callable, List, Optional, Tuple

import torch

from . import _native as N
from . import _tensor as T

_streams = {}
_lock = threading.RLock()  # guards the module-level caches (side streams, staging slots, captured graphs)


def _side_streams(dev: torch.device):
    key = (dev.type, dev.index)
    if key not in _streams:
        _streams[key] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
    return _streams[key]


_NSLOT = 3
_stage_cache = {}


def _staging(dev: torch.device, dtype, shapes):
    """Persistent device staging buffers (q, k, v chunk) x _NSLOT, cached per chunk geometry, so the pipelined loop
    never goes to the allocator (a cudaMalloc in the loop shows up as a multi-millisecond stall)."""
    key = (dev.index, dtype, shapes)
    slots = _stage_cache.get(key)
    if slots is None:
        if len(_stage_cache) >= 4:  # a handful of geometries at most: drop the oldest
            _stage_cache.pop(next(iter(_stage_cache)))
        slots = [tuple(torch.empty(sh, dtype=dtype, device=dev) for sh in shapes) for _ in range(_NSLOT)]
        _stage_cache[key] = slots
    return slots


def plan_chunks(B: int, Hq: int, Hkv: int, tensor_layout: str, chunks: Optional[int]) -> List[Tuple[int, int, int]]:
    """Cut the (batch, kv-head) units into chunks whose host memory is ONE contiguous range per tensor (a strided
    host slice would need a CPU-side gather before the DMA).  Returns [(b, kv_head_begin, kv_head_end), ...].
    HND `[B,H,N,D]`: any head range of one batch entry is contiguous.  NHD `[B,N,H,D]`: only whole batch entries are.
    `chunks` is a target count (default 8); q heads follow their kv head (GQA groups are never split).
    The pipeline is bound by the copy-in (PCIe), so what is left exposed is the tail: the last chunk's kernels and
    copy-out start only when its last input byte has arrived.  The last chunk is therefore cut once more into a large
    and a small part (a quarter), which shortens that tail without making the other chunks too small to fill the GPU."""
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
    plan = [(b, h0, h0 + step) for b in range(B) for h0 in range(0, Hkv, step)]
    if tensor_layout == "HND" and step >= 4 and len(plan) > 1:
        b, h0, h1 = plan.pop()
        cut = h1 - max(1, step // 4)
        plan += [(b, h0, cut), (b, cut, h1)]
    return plan


def _pipeline(qt, kt, vt, out, op, tensor_layout, plan, grp, dev, op_kwargs, keep=None, lse_out=None):
    """Enqueue the three-stream pipeline on the caller's current stream of `dev` (eagerly, or into a CUDA graph that
    is being captured on that stream: every event waited on is recorded inside this call and both side streams
    fork from and re-join the current stream, which is what stream capture requires).  `keep`: a list that receives
    the per-chunk outputs, so that under capture no chunk's memory is handed to a later chunk."""
    def view(t, b, h0, h1):
        return t[b:b + 1, h0:h1] if tensor_layout == "HND" else t[b:b + 1, :, h0:h1]

    cur = torch.cuda.current_stream(dev)
    s_in, s_out = _side_streams(dev)
    # everything enqueued so far on the caller's stream -- including an earlier call's kernels that read the staging
    # slots -- is ordered before the first copy of this call
    s_in.wait_stream(cur)
    s_out.wait_stream(cur)
    hmax = max(h1 - h0 for _, h0, h1 in plan)  # staging slots take the largest chunk; smaller ones use a leading view
    b0 = plan[0][0]
    shapes = tuple(tuple(view(t, b0, 0, e).shape) for t, e in ((qt, hmax * grp), (kt, hmax), (vt, hmax)))
    slots = _staging(dev, qt.dtype, shapes)
    if keep is not None:
        keep.append(slots)  # a captured graph writes into these buffers on every replay: it must own a reference,
        #                     or an eviction from the staging cache would hand their memory to someone else

    def lead(t, nh):  # the first nh heads of a staging buffer (contiguous: HND slots hold one batch entry)
        return t if tensor_layout != "HND" or t.shape[1] == nh else t[:, :nh]
    free = [None] * _NSLOT  # "the kernels that read this slot are done", recorded on the caller's stream

    def stage(i):
        """H2D of chunk i into staging slot i % NSLOT, after the kernels that last read that slot."""
        b, h0, h1 = plan[i]
        sq, sk, sv = slots[i % _NSLOT]
        dq, dk, dv = lead(sq, (h1 - h0) * grp), lead(sk, h1 - h0), lead(sv, h1 - h0)
        with torch.cuda.stream(s_in):
            if free[i % _NSLOT] is not None:
                s_in.wait_event(free[i % _NSLOT])
            dk.copy_(view(kt, b, h0, h1), non_blocking=True)  # K first: its mean + codes head the chunk
            dq.copy_(view(qt, b, h0 * grp, h1 * grp), non_blocking=True)
            dv.copy_(view(vt, b, h0, h1), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s_in)
        return dq, dk, dv, ev

    staged = stage(0)
    for i, (b, h0, h1) in enumerate(plan):
        torch.cuda.nvtx.range_push(f"lowbit.host chunk {i} (b {b}, kv heads {h0}:{h1})")  # host-side enqueue of the chunk
        dq, dk, dv, ev = staged
        if i + 1 < len(plan):
            staged = stage(i + 1)  # enqueue the next H2D before this chunk's kernels
        cur.wait_event(ev)
        o = op(dq, dk, dv, tensor_layout=tensor_layout, **op_kwargs)
        lse = None
        if lse_out is not None:  # return_lse=True: (o, lse [1, q heads of the chunk, N] fp32)
            o, lse = o
        done = torch.cuda.Event()
        done.record(cur)  # the slot may be overwritten once these kernels are done
        free[i % _NSLOT] = done
        with torch.cuda.stream(s_out):
            s_out.wait_event(done)
            view(out, b, h0 * grp, h1 * grp).copy_(o, non_blocking=True)
            if lse is not None:
                lse_out[b:b + 1, h0 * grp:h1 * grp].copy_(lse, non_blocking=True)
        if keep is not None:
            keep.append(o)
            keep.append(lse)
        else:
            o.record_stream(s_out)
            if lse is not None:
                lse.record_stream(s_out)
        torch.cuda.nvtx.range_pop()
    cur.wait_stream(s_in)
    cur.wait_stream(s_out)


_graphs = {}  # call signature -> [times seen, CUDAGraph | None, kept tensors]
_GRAPH_CACHE = 4


def lowbit_fa_host(q, k, v, out=None, op: Optional[Callable[..., Any]] = None, tensor_layout: str = "HND",
                   chunks: Optional[int] = None, device=None, graph: bool = False, lse_out=None, **op_kwargs: Any):
    """Run `op` (default `lowbit_fa_qk_int8_pv_fp16_triton`) on HOST tensors q, k, v (pinned memory for asynchronous
    DMA) and return the HOST tensor `out` (allocated pinned when not given), overlapping the copies with the kernels.
    Work is ordered on the caller's current stream of `device`: when this returns, everything is enqueued and the
    current stream has been made to wait for the last copy-out -- synchronize it (or an event on it) before reading
    `out` on the host.  With `return_lse=True` the call returns `(out, lse)`, lse `[B, Hq, N]` fp32 on the host
    (`lse_out`, allocated pinned when not given), copied out chunk by chunk like `out`.

    `graph=True` (opt-in): a caller that comes back with the SAME pinned buffers (a serving loop that refills them in
    place) gets the whole pipeline -- every copy, kernel and cross-stream dependency of every chunk -- as one CUDA
    graph, captured the second time that call signature is seen and replayed from then on.  The host then spends
    microseconds per call instead of ~0.2 ms per chunk.  What it costs, and why it is not the default: the capture
    synchronises the device once; every cached graph (up to 4 signatures, `drop_graphs()` releases them) keeps its
    staging slots, every chunk's output and all quantizer temporaries alive in a private pool (about one device copy
    of q, k, v, o plus the codes); and an `op` whose host-side control flow depends on data or on the environment is
    frozen as captured.  The default enqueues eagerly."""
    from . import core
    op = op or core.lowbit_fa_qk_int8_pv_fp16_triton
    want_lse = bool(op_kwargs.get("return_lse"))
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
    if want_lse:
        n_q = qt.shape[2] if tensor_layout == "HND" else qt.shape[1]
        if lse_out is None:
            lse_out = torch.empty((B, Hq, n_q), dtype=torch.float32, pin_memory=True)
        assert tuple(lse_out.shape) == (B, Hq, n_q) and lse_out.dtype == torch.float32 and lse_out.device.type == "cpu"
    else:
        lse_out = None
    ret = (out, lse_out) if want_lse else out

    key = None
    if graph and all(t.is_pinned() for t in (qt, kt, vt, out)) and (lse_out is None or lse_out.is_pinned()):
        try:
            key = (dev.index, qt.data_ptr(), kt.data_ptr(), vt.data_ptr(), out.data_ptr(),
                   lse_out.data_ptr() if lse_out is not None else 0, tuple(qt.shape),
                   tuple(kt.shape), tuple(vt.shape), qt.dtype, tensor_layout, len(plan), op,
                   tuple(sorted(op_kwargs.items())))
            hash(key)
        except TypeError:
            key = None  # an unhashable operator argument: enqueue eagerly
    with _lock, torch.cuda.device(dev):
        if key is None:
            _pipeline(qt, kt, vt, out, op, tensor_layout, plan, grp, dev, op_kwargs, lse_out=lse_out)
            return ret
        entry = _graphs.get(key)
        if entry is None:
            if len(_graphs) >= _GRAPH_CACHE:
                _graphs.pop(next(iter(_graphs)))
            _graphs[key] = [1, None, None]
            _pipeline(qt, kt, vt, out, op, tensor_layout, plan, grp, dev, op_kwargs, lse_out=lse_out)  # first sighting: eager (warms the kernels)
            return ret
        entry[0] += 1
        if entry[1] is None:
            cur = torch.cuda.current_stream(dev)
            cap = torch.cuda.Stream(dev)
            cap.wait_stream(cur)
            g, keep = torch.cuda.CUDAGraph(), []
            try:
                # thread_local: CUDA calls of other threads (e.g. a NCCL watchdog polling events) do not break the capture
                with torch.cuda.graph(g, stream=cap, capture_error_mode="thread_local"):
                    _pipeline(qt, kt, vt, out, op, tensor_layout, plan, grp, dev, op_kwargs, keep=keep, lse_out=lse_out)
                entry[1], entry[2] = g, keep
            except RuntimeError:
                entry[1] = False  # this operator cannot be captured (it synchronises): same kernels, enqueued eagerly
        if entry[1] is False:
            _pipeline(qt, kt, vt, out, op, tensor_layout, plan, grp, dev, op_kwargs, lse_out=lse_out)
        else:
            entry[1].replay()
    return ret


def drop_graphs() -> None:
    """Forget the captured pipelines (and release the device memory their private pools hold)."""
    with _lock:
        _graphs.clear()
