"""The framework hand-off (`_tensor.py`): Paddle is the reference's host framework and is not installed in this image,
so its branch is exercised with a stand-in that has exactly the surface `_tensor.py` touches -- `paddle.Tensor`,
`paddle.utils.dlpack.{to,from}_dlpack`, `paddle.device.cuda.current_stream().cuda_stream` -- over DLPack capsules of
real device memory.  What is checked: zero-copy views in both directions, results come back in the caller's framework,
and (GPU) the kernels run on the CALLER's stream, not on torch's current one.
"""
import types

import numpy as np
import pytest
import torch

from lowbit_quant_fa2_paddle_b200 import _tensor as T


class FakePaddleTensor:
    def __init__(self, t):
        self._t = t

    @property
    def shape(self):
        return list(self._t.shape)


def make_fake_paddle(stream_ptr=0):
    mod = types.SimpleNamespace()
    mod.Tensor = FakePaddleTensor
    mod.utils = types.SimpleNamespace(dlpack=types.SimpleNamespace(
        to_dlpack=lambda x: torch.utils.dlpack.to_dlpack(x._t),
        from_dlpack=lambda cap: FakePaddleTensor(torch.utils.dlpack.from_dlpack(cap))))
    mod.device = types.SimpleNamespace(cuda=types.SimpleNamespace(
        current_stream=lambda: types.SimpleNamespace(cuda_stream=stream_ptr)))
    return mod


def test_paddle_branch_is_zero_copy_both_ways(monkeypatch):
    monkeypatch.setattr(T, "paddle", make_fake_paddle())
    x = torch.arange(24, dtype=torch.float16).reshape(2, 3, 4)
    px = FakePaddleTensor(x)
    assert T.is_paddle(px) and not T.is_paddle(x)
    v = T.as_torch(px)
    assert isinstance(v, torch.Tensor) and v.data_ptr() == x.data_ptr() and v.shape == x.shape and v.stride() == x.stride()
    r = torch.ones(5)
    back = T.like(r, px)
    assert isinstance(back, FakePaddleTensor) and back._t.data_ptr() == r.data_ptr()
    assert T.like(r, x) is r and T.like(None, px) is None


def test_generic_dlpack_and_rejects():
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    v = T.as_torch(a)  # anything with __dlpack__
    assert isinstance(v, torch.Tensor) and v.shape == (2, 3) and float(v[1, 2]) == 5.0
    with pytest.raises(TypeError):
        T.as_torch([1, 2, 3])
    with pytest.raises(Exception):
        T.require_cuda(torch.zeros(2))  # no CPU fallback


def test_decorator_leaves_torch_callers_alone(monkeypatch):
    calls = []
    f = T.on_callers_stream(lambda *a, **k: calls.append((a, k)) or "ok")
    assert f(torch.zeros(1), x=2) == "ok" and len(calls) == 1
    monkeypatch.setattr(T, "paddle", make_fake_paddle(0))
    assert f(FakePaddleTensor(torch.zeros(1))) == "ok"  # CPU tensor / default stream: plain call


@pytest.mark.gpu
def test_paddle_caller_runs_on_paddles_stream(monkeypatch):
    import lowbit_quant_fa2_paddle_b200 as L
    dev = torch.device("cuda:0")
    side = torch.cuda.Stream(dev)
    monkeypatch.setattr(T, "paddle", make_fake_paddle(side.cuda_stream))
    torch.manual_seed(0)
    q, k, v = (torch.randn(1, 2, 512, 64, dtype=torch.float16, device=dev) for _ in range(3))
    ref = L.lowbit_fa_qk_int8_pv_fp16_triton(q, k, v)
    torch.cuda.synchronize(dev)
    seen = []
    orig = T.stream_ptr
    monkeypatch.setattr(T, "stream_ptr", lambda d: seen.append(orig(d)) or seen[-1])
    side.wait_stream(torch.cuda.current_stream(dev))
    out = L.lowbit_fa_qk_int8_pv_fp16_triton(FakePaddleTensor(q), FakePaddleTensor(k), FakePaddleTensor(v))
    side.synchronize()
    assert isinstance(out, FakePaddleTensor), "the result must come back in the caller's framework"
    # the operator forks one internal side stream off the caller's stream for the Q quantizer and joins it again; every
    # other launch -- the first and the last in particular -- must be on the caller's (Paddle's) stream, and none on the
    # stream torch considers current outside the call
    outside = torch.cuda.current_stream(dev).cuda_stream
    assert seen and seen[0] == side.cuda_stream and seen[-1] == side.cuda_stream and outside not in seen, \
        "kernels were not launched on the caller's (Paddle's) stream"
    assert sum(p == side.cuda_stream for p in seen) >= len(seen) - 1
    assert torch.equal(out._t, ref)
    km = L.k_mean(FakePaddleTensor(k))
    side.synchronize()
    assert isinstance(km, FakePaddleTensor) and torch.equal(km._t, L.k_mean(k))
