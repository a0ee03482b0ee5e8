#!/usr/bin/env python
"""Attention-kernel-only timing for a few shapes (development aid). usage: time_attn.py [shape ...]"""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import lowbit_quant_fa2_paddle_b200 as L
shapes = {"c2": (4, 32, 4096, 64, False), "c2c": (4, 32, 4096, 64, True), "d128": (4, 32, 4096, 128, False),
          "d128c8k": (4, 32, 8192, 128, True), "c2_16k": (1, 32, 16384, 64, False)}
dev = torch.device("cuda:0")
for name in (sys.argv[1:] or ["c2", "c2c", "d128", "d128c8k"]):
    b, h, n, d, causal = shapes[name]
    torch.manual_seed(0)
    q, k, v = (torch.randn(b, h, n, d, dtype=torch.float16, device=dev) for _ in range(3))
    km = L.k_mean(k)
    qi, qs, ki, ks = L.per_block_int8(q, k, km=km)
    f = L.forward_causal if causal else L.forward
    for _ in range(5):
        f(qi, ki, v, qs, ks)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20):
        f(qi, ki, v, qs, ks)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    ops = 4 * b * h * n * n * d / (2 if causal else 1)
    print(f"{name}: {ms:.3f} ms  {ops / ms / 1e9:.1f} TOPS  (variant {os.environ.get('LOWBIT_ATTN_VARIANT', '0')})", flush=True)
