#!/usr/bin/env python
"""HBM-roofline timing of the quantize kernels (development aid).  Rotates over several input tensors so that
no launch finds its input in the 126 MB L2.  usage: time_quant.py [c2|c3|c4]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import lowbit_quant_fa2_paddle_b200 as L  # noqa: E402
from lowbit_quant_fa2_paddle_b200 import _native as NV  # noqa: E402
from lowbit_quant_fa2_paddle_b200 import quant as Qz  # noqa: E402

shapes = {"c2": (4, 32, 4096, 64, "HND"), "c3": (4, 32, 8192, 128, "HND"), "c4": (2, 48, 17776, 64, "NHD")}
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
dev = torch.device("cuda:0")
for name in (sys.argv[1:] or ["c2", "c3"]):
    b, h, n, d, layout = shapes[name]
    shp = (b, h, n, d) if layout == "HND" else (b, n, h, d)
    xs = [torch.randn(shp, dtype=torch.float16, device=dev) for _ in range(4)]
    km = L.k_mean(xs[0], layout)
    elem = b * h * n * d

    def x2(x):  # a different tensor than x (the Q of the pair)
        for i, t in enumerate(xs):
            if t is x:
                return xs[(i + 2) % 4]
        return xs[0]

    def timeit(fn, reps=20):
        for i in range(4):
            fn(xs[i % 4])
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for i in range(reps):
            fn(xs[i % 4])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3

    cases = [
        ("k_mean", lambda x: L.k_mean(x, layout), 2.0),
        ("quant Q int8 (blk 128)", lambda x: Qz._quant_one(x, None, 128, 8, False, 0.18, NV.QMODE_TRITON, layout), 3.0),
        ("quant K int8 (blk 64, km)", lambda x: Qz._quant_one(x, km, 64, 8, False, 1.0, NV.QMODE_TRITON, layout), 3.0),
        ("quant K int4 packed", lambda x: Qz._quant_one(x, km, 64, 4, True, 1.0, NV.QMODE_TRITON, layout), 2.5),
        ("quant K int8 cuda-mode", lambda x: Qz._quant_one(x, km, 64, 8, False, 1.0, NV.QMODE_CUDA, layout), 3.0),
        ("V -> fp8 per channel", lambda x: L.per_channel_fp8(x, layout, smooth_v=False), 3.0),
        ("V -> fp8 per channel, smooth", lambda x: L.per_channel_fp8(x, layout, smooth_v=True), 3.0),
    ]
    # fused preparation: Q and K of the same shape (rotating pairs): 2 x (2 B read + 1 B written) per element of one
    # tensor = 6 B/elem algorithmic (K crosses HBM once; its second read is an L2 hit)
    if L.k_smooth_quant_supported(xs[0], layout):
        cases.append(("K fused (cluster): mean + K int8", lambda x: L.k_smooth_quant(x, 8, False, layout), 3.0))
        cases.append(("K separate: k_mean + K int8", lambda x: Qz._quant_one(x, L.k_mean(x, layout), 64, 8, False, 1.0, NV.QMODE_TRITON, layout), 3.0))
    cases.append(("unfused: k_mean + K + Q", lambda x: L.per_block_int8(x2(x), x, km=L.k_mean(x, layout), tensor_layout=layout), 6.0))
    for label, fn, bpe in cases:
        t = timeit(fn)
        gbs = elem * bpe / t / 1e9
        print(f"{name} {label:28s} {t * 1e6:8.1f} us  {gbs:7.0f} GB/s algorithmic ({bpe} B/elem)  {gbs / peak:5.2f} of measured HBM peak"
              f"  [tma={os.environ.get('LOWBIT_QUANT_TMA', '1')}]", flush=True)
