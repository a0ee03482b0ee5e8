#!/usr/bin/env python
"""Small driver for ncu: a few hot-path steps of one workload (default C2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import lowbit_quant_fa2_paddle_b200 as L  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = sys.argv[3] if len(sys.argv) > 3 else "i8f16"  # i8f16 | k4f16 | k4f8
shapes = {"c2": (4, 32, 4096, 64, False), "c2c": (4, 32, 4096, 64, True), "d128": (2, 16, 4096, 128, False),
          "d128c": (2, 16, 8192, 128, True)}
b, h, n, d, causal = shapes[wl]
torch.manual_seed(0)
dev = torch.device("cuda:0")
q, k, v = (torch.randn(b, h, n, d, dtype=torch.float16, device=dev) for _ in range(3))
fn = {"i8f16": L.lowbit_fa_qk_int8_pv_fp16_triton, "k4f16": L.lowbit_fa_qk_int4_pv_fp16_triton,
      "k4f8": L.lowbit_fa_qk_int4_pv_fp8}[mode]
for _ in range(steps):
    o = fn(q, k, v, is_causal=causal)
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
