#!/usr/bin/env python
"""End-to-end timing of lowbit_fa_host (pinned host q,k,v -> pinned host o) at BASELINE config 2 for several chunk
counts, eager and as a replayed CUDA graph (development aid).  usage: time_host.py [chunks ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import lowbit_quant_fa2_paddle_b200 as L  # noqa: E402

try:  # same CPU binding as bench.py: pinned buffers next to the GPU
    import bench
    bench.bind_near_gpu(0)
except Exception as e:  # noqa: BLE001
    print("no cpu binding:", e)
dev = torch.device("cuda:0")
b, h, n, d = 4, 32, 4096, 64
torch.manual_seed(0)
q, k, v = (torch.randn(b, h, n, d, dtype=torch.float16).pin_memory() for _ in range(3))
out = torch.empty_like(q).pin_memory()
ref = L.lowbit_fa_qk_int8_pv_fp16_triton(q.to(dev), k.to(dev), v.to(dev)).cpu()
ops = 4 * b * h * n * n * d
stream = torch.cuda.current_stream(dev)
for chunks in [int(a) for a in sys.argv[1:]] or [8, 16, 32, 64]:
    for graph in (False, True):
        out.zero_()
        for _ in range(3):
            L.lowbit_fa_host(q, k, v, out=out, chunks=chunks, graph=graph)
        torch.cuda.synchronize()
        assert torch.equal(out, ref), "host pipeline differs from the plain call"
        K = 20
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        ev[0].record(stream)
        for i in range(K):
            L.lowbit_fa_host(q, k, v, out=out, chunks=chunks, graph=graph)
            ev[i + 1].record(stream)
        torch.cuda.synchronize()
        per = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(K))
        med = per[K // 2]
        print(f"chunks={chunks:3d} graph={'replay' if graph is None else 'off   '}: median {med:.3f} ms  min {per[0]:.3f}  "
              f"max {per[-1]:.3f}   {ops / med / 1e9:.1f} TOPS", flush=True)
