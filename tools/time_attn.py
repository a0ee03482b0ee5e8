#!/usr/bin/env python
"""Attention-kernel-only timing for a few shapes and kernel modes (development aid).
usage: time_attn.py [shape[:mode] ...]   mode in i8f16 (default) | k4f16 | i8f8 | k4f8"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import lowbit_quant_fa2_paddle_b200 as L  # noqa: E402
from lowbit_quant_fa2_paddle_b200 import _native as NV  # noqa: E402
from lowbit_quant_fa2_paddle_b200 import attention as A  # noqa: E402

shapes = {"c2": (4, 32, 4096, 64, False), "c2c": (4, 32, 4096, 64, True), "d128": (4, 32, 4096, 128, False),
          "d128c8k": (4, 32, 8192, 128, True), "d128c16k": (2, 32, 16384, 128, True),
          "d128c32k": (1, 32, 32768, 128, True), "c2_16k": (1, 32, 16384, 64, False),
          "c4": (2, 48, 17776, 64, False), "d128_8k": (1, 32, 8192, 128, False)}
dev = torch.device("cuda:0")
for spec in (sys.argv[1:] or ["c2", "c2c", "d128", "d128c8k"]):
    name, _, mode = spec.partition(":")
    mode = mode or "i8f16"
    b, h, n, d, causal = shapes[name]
    torch.manual_seed(0)
    q, k, v = (torch.randn(b, h, n, d, dtype=torch.float16, device=dev) for _ in range(3))
    km = L.k_mean(k)
    if mode.startswith("k4"):
        qi, qs, ki, ks = L.per_block_q_int8_k_int4(q, k, km=km, pack=True)
        qk_mode = NV.QK_Q8K4
    else:
        qi, qs, ki, ks = L.per_block_int8(q, k, km=km)
        qk_mode = NV.QK_I8
    vs = vm = None
    pv_mode = NV.PV_F16
    if mode.endswith("f8"):
        v, vs, vm = L.per_channel_fp8(v, smooth_v=False)
        pv_mode = NV.PV_E4M3
    f = lambda: A._forward(qi, ki, v, qs, ks, "HND", torch.float16, False, causal, qk_mode=qk_mode, pv_mode=pv_mode,
                           v_scale=vs, v_mean=vm)
    for _ in range(5):
        f()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    reps = 20 if n <= 8192 else 5
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ops = 4 * b * h * n * n * d / (2 if causal else 1)
    print(f"{name}:{mode}: {ms:.3f} ms  {ops / ms / 1e9:.1f} TOPS  (N64={os.environ.get('LOWBIT_ATTN_N64', '1')} PF={os.environ.get('LOWBIT_ATTN_PF', 'default')})",
          flush=True)
