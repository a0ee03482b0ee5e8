#!/usr/bin/env python
"""Decode-shaped timing of the KV-cache attention path (csrc/kv_attn.cu): one query row per (batch, head) against a
KIVI-packed cache; reports achieved GB/s over the algorithmic bytes (codes + scales + minima, read once).
usage: time_kv.py [B H N D bits] ..."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from lowbit_quant_fa2_paddle_b200 import kv_cache as KV  # noqa: E402

try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6537.0
dev = torch.device("cuda:0")
cases = [(4, 32, 16384, 128, 4), (4, 32, 16384, 128, 2), (8, 32, 8192, 64, 4), (1, 32, 131072, 128, 4)]
if len(sys.argv) > 1:
    a = [int(x) for x in sys.argv[1:]]
    cases = [tuple(a[i:i + 5]) for i in range(0, len(a), 5)]
for (B, H, N, D, bits) in cases:
    torch.manual_seed(0)
    caches = []
    for _ in range(3):  # rotate over caches so that no launch finds its input in L2
        k = torch.randn(B, N, H, D, dtype=torch.float16, device=dev)
        v = torch.randn(B, N, H, D, dtype=torch.float16, device=dev)
        caches.append(KV.quant_and_pack_kv(k, v, 32, bits))
        del k, v
    q = torch.randn(B, 1, H, D, dtype=torch.float16, device=dev)
    nbytes = sum(t.numel() * t.element_size() for t in caches[0])
    for i in range(3):
        KV.quantized_flash_attn_forward(q, *caches[i], group_size=32, bits=bits)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    reps = 12
    e0.record()
    for i in range(reps):
        KV.quantized_flash_attn_forward(q, *caches[i % 3], group_size=32, bits=bits)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / reps * 1e-3
    print(f"B{B} H{H} N{N} D{D} {bits}-bit cache {nbytes / 1e6:7.1f} MB: {t * 1e6:8.1f} us  {nbytes / t / 1e9:7.0f} GB/s "
          f"({nbytes / t / 1e9 / peak:4.2f} of measured HBM peak)", flush=True)
