#!/usr/bin/env python
"""Latency of the operator at BASELINE config 1 (B1 H2 N512 D64, the reference's CPU-runnable case) and a few other
small shapes: eager back-to-back calls (host-bound at this size), the same call replayed from a CUDA graph (device
time of the 5 launches), torch SDPA for scale.  usage: time_small.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import lowbit_quant_fa2_paddle_b200 as L  # noqa: E402

dev = torch.device("cuda:0")


def ev_time(fn, reps):
    for _ in range(10):
        fn()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3, (t1 - t0) / reps * 1e6


for (B, H, N, D) in [(1, 2, 512, 64), (1, 8, 1024, 64), (2, 16, 1024, 128)]:
    torch.manual_seed(0)
    q, k, v = (torch.randn(B, H, N, D, dtype=torch.float16, device=dev) for _ in range(3))
    f = lambda: L.lowbit_fa_qk_int8_pv_fp16_triton(q, k, v)
    dev_us, host_us = ev_time(f, 200)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        f()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            o = f()
    gr_us, _ = ev_time(g.replay, 200)
    sd_us, _ = ev_time(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), 200)
    print(f"B{B} H{H} N{N} D{D}: eager {dev_us:7.1f} us/call (host enqueue {host_us:6.1f} us), graph replay {gr_us:6.1f} us, "
          f"torch SDPA {sd_us:6.1f} us", flush=True)
