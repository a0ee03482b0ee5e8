#!/usr/bin/env python
"""Attention kernel alone at config 2 (non-causal and causal), 30 back-to-back launches (development aid for the
LOWBIT_ATTN_* switches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import lowbit_quant_fa2_paddle_b200 as L  # noqa: E402
from lowbit_quant_fa2_paddle_b200 import attention as A  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
q, k, v = (torch.randn(4, 32, 4096, 64, dtype=torch.float16, device=dev) for _ in range(3))
km = L.k_mean(k)
qi, qs, ki, ks = L.per_block_int8(q, k, km=km)
tag = " ".join(f"{n}={os.environ[n]}" for n in sorted(os.environ) if n.startswith("LOWBIT_"))
for causal in (False, True):
    f = lambda: A._forward(qi, ki, v, qs, ks, "HND", torch.float16, False, causal)
    for _ in range(5):
        f()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(30):
        f()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 30
    print(f"{tag} causal={causal}: {ms * 1e3:.1f} us  {4 * 4 * 32 * 4096 * 4096 * 64 / (2 if causal else 1) / ms / 1e9:.1f} TOPS", flush=True)
