#!/usr/bin/env python
"""What one rank of the 8-way head-sharded config 4 does, on ONE GPU: the kv-head slice [h0:h1] of the full NHD
tensors (B2 H48 N17776 D64, q8/k4).  Prints whole-operator and attention-only times for 48 / 24 / 12 / 6 heads, i.e.
the per-GPU work at 1 / 2 / 4 / 8 GPUs, so that strong-scaling losses can be studied without an 8-GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import lowbit_quant_fa2_paddle_b200 as L  # noqa: E402
from lowbit_quant_fa2_paddle_b200 import _native as NV  # noqa: E402
from lowbit_quant_fa2_paddle_b200 import attention as A  # noqa: E402
from lowbit_quant_fa2_paddle_b200 import parallel as P  # noqa: E402

dev = torch.device("cuda:0")
B, H, N, D = 2, 48, 17776, 64
torch.manual_seed(0)
q, k, v = (torch.randn(B, N, H, D, dtype=torch.float16, device=dev) for _ in range(3))


def timed(f, reps=10):
    for _ in range(3):
        f()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for world in (1, 2, 4, 8):
    hs = H // world
    ops = 4.0 * B * hs * N * N * D
    op = lambda: P.lowbit_fa_head_sharded(q, k, v, L.lowbit_fa_q_int8_k_int4_pv_fp16, world, 0, tensor_layout="NHD")  # noqa: E731
    t_op = timed(op)
    qs_, ks_, vs_ = q[:, :, :hs], k[:, :, :hs], v[:, :, :hs]
    km = L.k_mean(ks_, "NHD")
    qi, qsc, ki, ksc = L.per_block_q_int8_k_int4(qs_, ks_, km=km, tensor_layout="NHD")
    att = lambda: A._forward(qi, ki, vs_, qsc, ksc, "NHD", torch.float16, False, False, qk_mode=NV.QK_Q8K4)  # noqa: E731
    t_att = timed(att)
    qi8, qsc8, ki8, ksc8 = L.per_block_int8(qs_, ks_, km=km, tensor_layout="NHD")
    att8 = lambda: A._forward(qi8, ki8, vs_, qsc8, ksc8, "NHD", torch.float16, False, False)  # noqa: E731
    t_att8 = timed(att8)
    vc = vs_.contiguous()
    attc = lambda: A._forward(qi, ki, vc, qsc, ksc, "NHD", torch.float16, False, False, qk_mode=NV.QK_Q8K4)  # noqa: E731
    t_attc = timed(attc)
    ctas = ((N + 127) // 128) * hs * B
    print(f"1/{world} of the heads ({hs}): operator {t_op:.3f} ms ({ops / t_op / 1e9:.0f} TOPS)  attention q8k4 {t_att:.3f} ms "
          f"({ops / t_att / 1e9:.0f})  [contiguous V {t_attc:.3f} ms]  attention int8 {t_att8:.3f} ms ({ops / t_att8 / 1e9:.0f})  "
          f"CTAs {ctas} = {ctas / 592:.2f} waves of 592", flush=True)
