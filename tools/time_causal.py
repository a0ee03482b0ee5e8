#!/usr/bin/env python
"""Causal attention timing (development aid): quantize + attention through the public API, C2-causal and C3 @ 8K / 16K;
LOWBIT_CAUSAL_SECTION sets the (batch, head) section of the longest-first tile order."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import lowbit_quant_fa2_paddle_b200 as L  # noqa: E402

dev = torch.device("cuda:0")
cases = [("c2c  D64  N4096  int8/fp16", 4, 32, 4096, 64, L.lowbit_fa_qk_int8_pv_fp16_triton),
         ("c3   D128 N8192  int4/fp8 ", 4, 32, 8192, 128, L.lowbit_fa_qk_int4_pv_fp8),
         ("c3   D128 N16384 int4/fp8 ", 2, 32, 16384, 128, L.lowbit_fa_qk_int4_pv_fp8),
         ("     D128 N8192  int8/fp16", 4, 32, 8192, 128, L.lowbit_fa_qk_int8_pv_fp16_triton)]
for name, b, h, n, d, fn in cases:
    torch.manual_seed(0)
    q, k, v = (torch.randn(b, h, n, d, dtype=torch.float16, device=dev) for _ in range(3))
    f = lambda: fn(q, k, v, is_causal=True)
    for _ in range(3):
        f()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms:8.3f} ms  {2.0 * b * h * n * n * d / ms / 1e9:7.1f} TOPS  (section {os.environ.get('LOWBIT_CAUSAL_SECTION', '32')})", flush=True)
    del q, k, v
