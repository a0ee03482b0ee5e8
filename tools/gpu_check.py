#!/usr/bin/env python
"""Verbose on-GPU diagnostics (development aid; the real gates are tests/ -m gpu).
Prints quantizer exactness, the raw int32 score dump of one tile against an integer matmul, and attention
errors versus the CPU oracle for a sweep of shapes.  Usage: python tools/gpu_check.py [quick]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import lowbit_quant_fa2_paddle_b200 as L  # noqa: E402
from lowbit_quant_fa2_paddle_b200 import _native as N  # noqa: E402
from oracle import attention as OA  # noqa: E402
from oracle import quant as OQ  # noqa: E402

dev = torch.device("cuda:0")


def mk(b, h, n, d, layout, dtype, seed, bias=0.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, h, n, d, generator=g)
    if bias:
        x = x + bias * torch.randn(1, h, 1, d, generator=g)
    x = x.to(dtype)
    return x if layout == "HND" else x.permute(0, 2, 1, 3).contiguous()


def check_quant():
    print("== quantizers ==")
    for layout in ("HND", "NHD"):
        for dtype in (torch.float16, torch.bfloat16):
            for (b, h, n, d) in ((1, 2, 512, 64), (2, 3, 200, 128), (1, 1, 1, 64), (1, 2, 4096, 64)):
                q = mk(b, h, n, d, layout, dtype, 1)
                k = mk(b, h, n, d, layout, dtype, 2, bias=3.0)
                km = L.k_mean(k.to(dev), layout)
                km_ref = OQ.k_mean(k, layout)
                ok_km = torch.equal(km.cpu(), km_ref)
                for backend, ofn in (("triton", OQ.per_block_int8_q1), ("cuda", OQ.per_block_int8_q2)):
                    got = L.per_block_int8(q.to(dev), k.to(dev), km=km, tensor_layout=layout, backend=backend)
                    ref = ofn(q, k, km_ref, tensor_layout=layout)
                    res = [bool(torch.equal(g.cpu(), r)) for g, r in zip(got, ref)]
                    nd = [int((g.cpu() != r).sum()) for g, r in zip(got, ref)]
                    print(f"  {layout} {str(dtype)[6:]:8s} B{b} H{h} N{n} D{d} {backend:6s} km={ok_km} exact={res} ndiff={nd}")
                got = L.per_block_q_int8_k_int4(q.to(dev), k.to(dev), km=km, tensor_layout=layout)
                ref = OQ.per_block_int8_q1(q, k, km_ref, tensor_layout=layout, kbits=4)
                ok = torch.equal(got[2].cpu(), OQ.pack_codes(ref[2], 4)) and torch.equal(got[3].cpu(), ref[3])
                print(f"     packed int4 K exact={ok}")


def check_scores():
    print("== raw int32 score dump (tile 0, key block 0) ==")
    for d in (64, 128):
        q = mk(1, 1, 128, d, "HND", torch.float16, 3)
        k = mk(1, 1, 128, d, "HND", torch.float16, 4)
        v = mk(1, 1, 128, d, "HND", torch.float16, 5)
        qi, qs, ki, ks = OQ.per_block_int8_q1(q, k, None)
        buf = torch.zeros(128 * 64, dtype=torch.int32, device=dev)
        N.lib().lowbit_attn_set_debug_buffer(buf.data_ptr())
        o, _ = L.forward(qi.to(dev), ki.to(dev), v.to(dev), qs.to(dev), ks.to(dev))
        torch.cuda.synchronize()
        N.lib().lowbit_attn_set_debug_buffer(None)
        S = buf.cpu().view(128, 64)
        Sref = (qi[0, 0].int() @ ki[0, 0, :64].int().T)
        bad = (S != Sref)
        print(f"  D={d}: mismatches {int(bad.sum())}/{bad.numel()}")
        if bad.any():
            rows = bad.any(dim=1).nonzero().flatten().tolist()
            cols = bad.any(dim=0).nonzero().flatten().tolist()
            print("   bad rows:", rows[:40], "...")
            print("   bad cols:", cols[:40], "...")
            print("   S[0,:8]   ", S[0, :8].tolist())
            print("   Sref[0,:8]", Sref[0, :8].tolist())
            # is S a permutation of columns/rows of Sref?
            for r in range(0, 3):
                match = [(int((Sref[:, c] == S[:, r]).all())) for c in range(64)]
                print(f"   col {r} of S equals Sref col:", [i for i, mm in enumerate(match) if mm])


def check_attn(quick=False):
    print("== attention vs oracle ==")
    cases = [
        (1, 1, 1, 128, 64, "HND", False), (1, 1, 1, 128, 128, "HND", False),
        (1, 2, 2, 512, 64, "HND", False), (1, 2, 2, 512, 64, "HND", True),
        (1, 2, 2, 512, 128, "HND", False), (1, 2, 2, 512, 128, "HND", True),
        (1, 4, 2, 384, 64, "NHD", False), (2, 4, 2, 384, 128, "NHD", True),
        (1, 2, 2, 200, 64, "HND", False), (1, 2, 2, 200, 64, "HND", True), (1, 2, 1, 77, 128, "NHD", False),
        (1, 1, 1, 1, 64, "HND", False), (1, 2, 2, 1000, 64, "HND", True),
    ]
    if not quick:
        cases += [(1, 2, 2, 2048, 64, "HND", False), (1, 2, 2, 2048, 128, "HND", True)]
    for (b, hq, hkv, n, d, layout, causal) in cases:
        q = mk(b, hq, n, d, layout, torch.float16, 11)
        k = mk(b, hkv, n, d, layout, torch.float16, 12, bias=2.0)
        v = mk(b, hkv, n, d, layout, torch.float16, 13)
        t0 = time.time()
        o, lse = L.lowbit_fa_qk_int8_pv_fp16_triton(q.to(dev), k.to(dev), v.to(dev), tensor_layout=layout,
                                                    is_causal=causal, return_lse=True)
        torch.cuda.synchronize()
        t1 = time.time()
        oref, lref = OA.lowbit_fa_api(q, k, v, layout, causal, return_lse=True, compat_tail=False, pv_accum="fp32")
        sd = OA.sdpa_fp32(q, k, v, layout, causal)
        e = (o.cpu().float() - oref.float()).abs().max().item()
        cs = torch.nn.functional.cosine_similarity(o.cpu().float().flatten(), oref.float().flatten(), dim=0).item()
        cs2 = torch.nn.functional.cosine_similarity(o.cpu().float().flatten(), sd.flatten(), dim=0).item()
        le = (lse.cpu() - lref).abs().max().item()
        nan = int(torch.isnan(o.float()).sum())
        print(f"  B{b} Hq{hq} Hkv{hkv} N{n} D{d} {layout} causal={int(causal)}: max|o-oracle|={e:.2e} cos={cs:.6f} "
              f"cos_sdpa={cs2:.6f} lse_err={le:.2e} nan={nan} ({(t1 - t0) * 1e3:.1f} ms)")


def quick_bench():
    print("== quick timing (C2: B4 H32 N4096 D64) ==")
    b, h, n, d = 4, 32, 4096, 64
    q = torch.randn(b, h, n, d, dtype=torch.float16, device=dev)
    k = torch.randn(b, h, n, d, dtype=torch.float16, device=dev)
    v = torch.randn(b, h, n, d, dtype=torch.float16, device=dev)
    km = L.k_mean(k)
    qi, qs, ki, ks = L.per_block_int8(q, k, km=km)
    for causal in (False, True):
        f = L.forward_causal if causal else L.forward
        for _ in range(3):
            f(qi, ki, v, qs, ks)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            f(qi, ki, v, qs, ks)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        ops = 4 * b * h * n * n * d / (2 if causal else 1)
        print(f"  attention causal={int(causal)}: {ms:.3f} ms  {ops / ms / 1e9:.1f} TOPS")
    for name, fn in (("k_mean", lambda: L.k_mean(k)), ("quant q+k", lambda: L.per_block_int8(q, k, km=km))):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"  {name}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")


if __name__ == "__main__":
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    print(torch.cuda.get_device_name(0))
    check_quant()
    check_scores()
    check_attn(quick)
    quick_bench()
