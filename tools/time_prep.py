#!/usr/bin/env python
"""Where the step time outside the attention kernel goes at config 2 (development aid): full operator, K chain +
attention (Q codes given), attention only, K chain only, Q quantizer only; back-to-back launches, CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import lowbit_quant_fa2_paddle_b200 as L  # noqa: E402
from lowbit_quant_fa2_paddle_b200 import _native as NV  # noqa: E402
from lowbit_quant_fa2_paddle_b200 import attention as A  # noqa: E402
from lowbit_quant_fa2_paddle_b200 import quant as Qz  # noqa: E402

b, h, n, d = 4, 32, 4096, 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
q, k, v = (torch.randn(b, h, n, d, dtype=torch.float16, device=dev) for _ in range(3))
km = L.k_mean(k)
qi, qs, ki, ks = L.per_block_int8(q, k, km=km)
sc = d ** -0.5 * 1.4426950408889634


def t(fn, reps=30):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def kchain():
    m = L.k_mean(k)
    return Qz._quant_one(k, m, 64, 8, False, 1.0, NV.QMODE_TRITON, "HND")


def kchain_attn():
    kc, ks2 = kchain()
    return A._forward(qi, kc, v, qs, ks2, "HND", torch.float16, False, False)


cases = [("full operator", lambda: L.lowbit_fa_qk_int8_pv_fp16_triton(q, k, v)),
         ("K chain + attention", kchain_attn),
         ("attention only", lambda: A._forward(qi, ki, v, qs, ks, "HND", torch.float16, False, False)),
         ("K chain only", kchain),
         ("k_mean only", lambda: L.k_mean(k)),
         ("K quant only", lambda: Qz._quant_one(k, km, 64, 8, False, 1.0, NV.QMODE_TRITON, "HND")),
         ("Q quant only", lambda: Qz._quant_one(q, None, 128, 8, False, sc, NV.QMODE_TRITON, "HND"))]
for name, fn in cases:
    print(f"{name:24s} {t(fn):8.1f} us", flush=True)
