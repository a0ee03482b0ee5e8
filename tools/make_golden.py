#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the reference's OWN Triton kernels (from /root/reference,
unmodified) under the Triton CPU interpreter.  Run in the build container only:

    python tools/make_golden.py

The fixtures are committed; the GPU box and the test-suite never need /root/reference.
fp16/bf16 tensors are stored as their uint16 bit patterns (npz has no bf16)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import ref_triton as R  # noqa: E402
from oracle import quant as Q  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def bits16(t):
    return t.contiguous().view(torch.int16).numpy().view(np.uint16)


def mk(shape_hnd, layout, dtype, seed, bias=0.0):
    g = torch.Generator().manual_seed(seed)
    b, h, n, d = shape_hnd
    x = torch.randn(b, h, n, d, generator=g)
    if bias:
        x = x + bias * torch.randn(1, h, 1, d, generator=g)
    x = x.to(dtype)
    return x if layout == "HND" else x.permute(0, 2, 1, 3).contiguous()


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            if v.dtype in (torch.float16, torch.bfloat16):
                out[k + "__" + str(v.dtype).split(".")[-1]] = bits16(v)
            else:
                out[k] = v.contiguous().numpy()
        else:
            out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, {k: v.shape for k, v in out.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    # ---- quantizers (Q1 int8, Q4 int4-unpack, Q3 per-thread) + attention (A1/A2) ----
    cases = [
        # name, B,Hq,Hkv,N,D, layout, dtype, causal
        ("c1_hnd_d64", 1, 2, 2, 512, 64, "HND", torch.float16, False),   # BASELINE config 1
        ("c1_hnd_d64_causal", 1, 2, 2, 512, 64, "HND", torch.float16, True),
        ("nhd_d128_gqa", 1, 4, 2, 256, 128, "NHD", torch.float16, False),
        ("nhd_d128_gqa_causal", 1, 4, 2, 256, 128, "NHD", torch.float16, True),
        ("tail_d64", 1, 2, 2, 200, 64, "HND", torch.float16, False),     # Nk % 64 != 0: phantom keys
        ("tail_d64_causal", 1, 2, 2, 200, 64, "HND", torch.float16, True),
        ("bf16_d64", 1, 2, 1, 192, 64, "HND", torch.bfloat16, False),
    ]
    for i, (name, b, hq, hkv, n, d, layout, dtype, causal) in enumerate(cases):
        q = mk((b, hq, n, d), layout, dtype, 100 + i)
        k = mk((b, hkv, n, d), layout, dtype, 200 + i, bias=3.0)
        v = mk((b, hkv, n, d), layout, dtype, 300 + i)
        km = Q.k_mean(k, layout)
        sm = d ** -0.5
        qi, qs, ki, ks = R.per_block_int8(q, k, km, sm_scale=sm, tensor_layout=layout)
        _, _, k4, k4s = R.per_block_int8(q, k, km, sm_scale=sm, tensor_layout=layout, kbits=4)
        v16 = v.to(torch.float16)
        o, lse2 = R.attn_forward(qi, ki, v16, qs, ks, layout, causal, output_dtype=dtype, return_lse=True)
        o4, _ = R.attn_forward(qi, k4, v16, qs, k4s, layout, causal, output_dtype=dtype, return_lse=False)
        extra = {}
        if not causal:
            tq, tqs, tk, tks = R.per_thread(q, k, km, tensor_layout=layout, bits=8)
            uq, uqs, uk, uks = R.per_thread(q, k, km, tensor_layout=layout, bits=4)
            extra = dict(pt8_q=tq, pt8_qs=tqs, pt8_k=tk, pt8_ks=tks, pt4_q=uq, pt4_qs=uqs, pt4_k=uk, pt4_ks=uks)
        save("attn_" + name, q=q, k=k, v=v, km=km, sm_scale=sm, layout=layout, causal=causal,
             q_int8=qi, q_scale=qs, k_int8=ki, k_scale=ks, k_int4=k4, k_int4_scale=k4s,
             o=o, lse2=lse2, o_k4=o4, **extra)
    # ---- KIVI asymmetric pack (Q5) ----
    for bit in (2, 4, 8):
        g = torch.Generator().manual_seed(900 + bit)
        data = torch.randn(1, 2, 16, 128, generator=g).to(torch.float16)
        code, scale, mn = R.kivi_quantize_and_pack(data, 32, bit)
        save(f"kivi_b{bit}", data=data, code=code, scale=scale, mn=mn, bit=bit, group_size=32)


if __name__ == "__main__":
    main()
