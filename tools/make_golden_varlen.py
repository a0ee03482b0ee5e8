#!/usr/bin/env python
"""Generate tests/golden/varlen_*.npz by running the reference's OWN varlen Triton kernels (quantizer, non-causal and
causal attention; from /root/reference, unmodified) under the Triton CPU interpreter.  Build container only:

    python tools/make_golden_varlen.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import ref_triton as R  # noqa: E402
from oracle import varlen as OV  # noqa: E402
from make_golden import save  # noqa: E402


def main():
    cases = [
        # name, Hq, Hkv, D, dtype, causal, q lengths, k lengths
        ("d64", 2, 2, 64, torch.float16, False, [200, 64, 333], [130, 257, 64]),
        ("d64_causal", 2, 2, 64, torch.float16, True, [200, 64, 333], [200, 64, 333]),
        ("d128_gqa", 4, 2, 128, torch.float16, False, [128, 70], [192, 100]),
        ("d128_gqa_causal_bf16", 4, 2, 128, torch.bfloat16, True, [128, 70], [128, 70]),
    ]
    for i, (name, hq, hkv, d, dtype, causal, lq, lk) in enumerate(cases):
        g = torch.Generator().manual_seed(700 + i)
        q = torch.randn(sum(lq), hq, d, generator=g).to(dtype)
        k = (torch.randn(sum(lk), hkv, d, generator=g) + 2.0 * torch.randn(1, hkv, d, generator=g)).to(dtype)
        v = torch.randn(sum(lk), hkv, d, generator=g).to(dtype)
        cu_q = torch.tensor([0] + list(torch.tensor(lq).cumsum(0)), dtype=torch.int32)
        cu_k = torch.tensor([0] + list(torch.tensor(lk).cumsum(0)), dtype=torch.int32)
        km = OV.k_mean_varlen(k)          # km contract (the reference's Paddle reduction order is unknowable)
        ks = k - km                       # core.py:449, in the input dtype
        sm = d ** -0.5
        qi, qs, ki, kss, cqs, cks = R.varlen_per_block_int8(q, ks, cu_q, cu_k, max(lq), max(lk), sm_scale=sm)
        v16 = v.to(torch.float16)
        o = R.varlen_attn_forward(qi, ki, v16, cu_q, cu_k, max(lq), qs, kss, cqs, cks, causal, dtype)
        save("varlen_" + name, q=q, k=k, v=v, km=km, cu_q=cu_q, cu_k=cu_k, sm_scale=sm, causal=causal,
             q_int8=qi, q_scale=qs, k_int8=ki, k_scale=kss, cu_q_scale=cqs, cu_k_scale=cks, o=o)


if __name__ == "__main__":
    main()
