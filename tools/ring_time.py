#!/usr/bin/env python
"""Time ring_attention for one (N, qk, pv) under torchrun (development aid).  usage: ring_time.py N qk pv [reps]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from lowbit_quant_fa2_paddle_b200 import parallel as P  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, qk, pv = int(sys.argv[1]), sys.argv[2], sys.argv[3]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
torch.manual_seed(rank)
q, k, v = (torch.randn(1, 32, N // world, 128, dtype=torch.float16, device=dev) for _ in range(3))
f = lambda: P.ring_attention(q, k, v, is_causal=True, qk=qk, pv=pv)  # noqa: E731
for _ in range(3):
    f()
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(reps):
    f()
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ops = 4.0 * 32 * N * N * 128 / 2
    print(f"ring world={world} N={N} qk={qk} pv={pv}: {t.item():.3f} ms  {ops / t.item() / 1e9:.0f} TOPS", flush=True)
dist.barrier()
dist.destroy_process_group()
