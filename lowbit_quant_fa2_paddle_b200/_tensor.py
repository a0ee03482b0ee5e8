"""Framework-neutral tensor hand-off.

The reference operates on Paddle tensors (`q.place`, `q.strides`, `paddle.empty`).  The CUDA library only
needs device pointers, shapes and element strides, so anything that speaks DLPack works: Paddle and torch
tensors are viewed zero-copy as torch tensors (torch is the device-memory/stream plumbing of this build),
outputs are allocated by torch's caching allocator and handed back to the caller's framework via DLPack.
"""
import torch

from . import _native as N

try:  # Paddle is optional; the reference's host framework
    import paddle  # type: ignore
    if getattr(paddle, "__lowbit_stub__", False):
        paddle = None
except Exception:  # pragma: no cover - paddle absent in the build container
    paddle = None


def is_paddle(t):
    return paddle is not None and isinstance(t, paddle.Tensor)


def as_torch(t):
    """Zero-copy torch view of a torch / Paddle / DLPack-capable tensor."""
    if isinstance(t, torch.Tensor):
        return t
    if is_paddle(t):
        return torch.utils.dlpack.from_dlpack(paddle.utils.dlpack.to_dlpack(t))
    if hasattr(t, "__dlpack__"):
        return torch.from_dlpack(t)
    raise TypeError(f"unsupported tensor type {type(t)!r}: need a torch / Paddle / DLPack tensor")


def like(result, original):
    """Return `result` (torch) in the framework of `original`."""
    if result is None or isinstance(original, torch.Tensor):
        return result
    if is_paddle(original):
        return paddle.utils.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(result))
    return result


def _paddle_stream_ptr():
    """Raw cudaStream_t of Paddle's current stream (0 = the legacy default stream), or None when it cannot be read."""
    try:
        return int(paddle.device.cuda.current_stream().cuda_stream)
    except Exception:
        try:
            return int(paddle.device.current_stream().stream_base.cuda_stream)
        except Exception:
            return None


def on_callers_stream(fn):
    """Decorator of the public entry points.  torch callers: untouched (kernels go to torch's current stream).  Paddle
    callers: the call runs with Paddle's CURRENT stream installed as torch's current stream (an ExternalStream), so the
    kernels, torch's caching allocator and the DLPack hand-back are all ordered on the stream the caller's own work
    is on -- no reliance on both frameworks happening to use the legacy default stream."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        if paddle is None:
            return fn(*args, **kwargs)
        first = next((a for a in list(args) + list(kwargs.values()) if is_paddle(a)), None)
        if first is None:
            return fn(*args, **kwargs)
        dev = as_torch(first).device
        ptr = _paddle_stream_ptr() if dev.type == "cuda" else None
        if not ptr:  # CPU tensor (the entry point raises), unreadable stream, or the legacy default stream (= torch's default)
            return fn(*args, **kwargs)
        with torch.cuda.stream(torch.cuda.ExternalStream(ptr, device=dev)):
            return fn(*args, **kwargs)
    return wrapped


def aligned16(t):
    """The kernels load 16 bytes at a time: a view whose storage offset breaks that alignment (e.g. x[..., 4:68] of a
    fused buffer -- legal for the reference's Triton kernels) is copied to a fresh, aligned tensor."""
    return t if t.data_ptr() % 16 == 0 else t.clone(memory_format=torch.contiguous_format)


def dtype_code(dt):
    if dt == torch.float16:
        return N.F16
    if dt == torch.bfloat16:
        return N.BF16
    raise AssertionError("Input tensors must be in dtype of torch.float16 or torch.bfloat16")


def require_cuda(*ts):
    dev = ts[0].device
    for t in ts:
        if t.device.type != "cuda":
            raise N.LowbitNativeError(
                "lowbit_fa: tensors must live on a CUDA device (sm_100a); there is no CPU fallback")
        assert t.device == dev, "All tensors must be on the same device."
    return dev


def stream_ptr(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def bhnd(t, layout):
    """(B, H, N, D, stride_b, stride_h, stride_n) of a 4-D tensor in the reference's HND / NHD layouts
    (quant_per_block.py:188-201)."""
    if layout == "HND":
        b, h, n, d = t.shape
        sb, sh, sn = t.stride(0), t.stride(1), t.stride(2)
    elif layout == "NHD":
        b, n, h, d = t.shape
        sb, sn, sh = t.stride(0), t.stride(1), t.stride(2)
    else:
        raise ValueError(f"Unknown tensor layout: {layout}")
    assert t.stride(3) == 1, "Last dim of qkv must be contiguous."
    return b, h, n, d, sb, sh, sn
