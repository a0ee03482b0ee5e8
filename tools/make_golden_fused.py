#!/usr/bin/env python
"""tools/make_golden_fused.py -- golden vectors for rows Q2 / Q6 from the reference's OWN CUDA kernels.

Runs on a B200 (the kernels are CUDA).  Imports the extension that oracle/build_ref_fused.py compiled from
/root/reference/csrc/fused/{fused.cu,pybind.cpp} (unmodified) and calls it exactly as src/quant.py does:

    per_block_int8   src/quant.py:70-98    quant_per_block_int8_cuda / ..._fuse_sub_mean_cuda
    per_warp_int8    src/quant.py:147-172  quant_per_warp_int8_cuda
    sub_mean         src/quant.py:203-208  sub_mean_cuda
    per_channel_fp8  src/quant.py:254-291  transpose_pad_permute_cuda + (mean_)scale_fuse_quant_cuda

Writes inputs + outputs as tests/golden/fused_*.npz (or --out DIR; on the GPU box: gpurun_out/golden_fused, copied
into tests/golden/ afterwards).  `--variant fast` runs the --use_fast_math build and only REPORTS how many codes
differ from the IEEE build (documented in DESIGN.md; not stored).
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref_fused as BRF  # noqa: E402

LOG2E = 1.44269504


def _bits(t):
    """fp16 / bf16 / fp8 tensors as integer bit patterns (npz has no such dtypes)."""
    if t.dtype in (torch.float16, torch.bfloat16):
        return t.view(torch.int16).cpu().numpy()
    if t.dtype == torch.float8_e4m3fn:
        return t.view(torch.uint8).cpu().numpy()
    return t.cpu().numpy()


def ref_per_block_int8(F, q, k, km, BLKQ, BLKK, sm_scale, layout):
    tl = 0 if layout == "NHD" else 1
    if layout == "HND":
        b, hq, nq, d = q.shape
        _, hkv, nk, _ = k.shape
    else:
        b, nq, hq, d = q.shape
        _, nk, hkv, _ = k.shape
    qi, ki = torch.empty_like(q, dtype=torch.int8), torch.empty_like(k, dtype=torch.int8)
    qs = torch.empty((b, hq, (nq + BLKQ - 1) // BLKQ), dtype=torch.float32, device=q.device)
    ks = torch.empty((b, hkv, (nk + BLKK - 1) // BLKK), dtype=torch.float32, device=q.device)
    F.quant_per_block_int8_cuda(q, qi, qs, float(sm_scale * LOG2E), BLKQ, tl)
    if km is not None:
        F.quant_per_block_int8_fuse_sub_mean_cuda(k, km.squeeze(1) if tl == 0 else km.squeeze(2), ki, ks, BLKK, tl)
    else:
        F.quant_per_block_int8_cuda(k, ki, ks, BLKK, tl)
    return qi, qs, ki, ks


def ref_per_warp_int8(F, q, k, km, layout):
    tl = 0 if layout == "NHD" else 1
    if layout == "HND":
        b, hq, nq, d = q.shape
        _, hkv, nk, _ = k.shape
    else:
        b, nq, hq, d = q.shape
        _, nk, hkv, _ = k.shape
    qi, ki = torch.empty_like(q, dtype=torch.int8), torch.empty_like(k, dtype=torch.int8)
    qs = torch.empty((b, hq, (nq + 127) // 128 * 4), dtype=torch.float32, device=q.device)
    ks = torch.empty((b, hkv, (nk + 63) // 64), dtype=torch.float32, device=q.device)
    F.quant_per_warp_int8_cuda(q, qi, qs, tl)
    if km is not None:
        F.quant_per_block_int8_fuse_sub_mean_cuda(k, km.squeeze(1) if tl == 0 else km.squeeze(2), ki, ks, 64, tl)
    else:
        F.quant_per_block_int8_cuda(k, ki, ks, 64, tl)
    return qi, qs, ki, ks


def ref_per_channel_fp8(F, v, layout, scale_max, smooth_v):
    tl = 0 if layout == "NHD" else 1
    if layout == "HND":
        b, h, n, d = v.shape
        vt = torch.empty((b, h, d, (n + 63) // 64 * 64), dtype=v.dtype, device=v.device)
    else:
        b, n, h, d = v.shape
        vt = torch.empty((b, d, h, (n + 63) // 64 * 64), dtype=v.dtype, device=v.device)
    F.transpose_pad_permute_cuda(v, vt, tl)
    v8 = torch.empty(vt.shape, dtype=torch.float8_e4m3fn, device=v.device)
    vs = torch.empty((b, h, d), dtype=torch.float32, device=v.device)
    vm = torch.empty((b, h, d), dtype=torch.float32, device=v.device)
    if smooth_v:
        F.mean_scale_fuse_quant_cuda(vt, v8, vm, vs, n, float(scale_max), tl)
        return vt, v8, vs, vm
    F.scale_fuse_quant_cuda(vt, v8, vs, n, float(scale_max), tl)
    return vt, v8, vs, None


CASES = [
    # name, B, Hq, Hkv, Nq, Nk, D, layout, dtype
    ("hnd_f16_d64", 1, 2, 2, 256, 320, 64, "HND", torch.float16),
    ("hnd_f16_d64_tail", 2, 4, 2, 200, 333, 64, "HND", torch.float16),
    ("nhd_f16_d128_tail", 1, 3, 3, 130, 257, 128, "NHD", torch.float16),
    ("hnd_bf16_d128", 1, 2, 1, 192, 128, 128, "HND", torch.bfloat16),
    ("nhd_bf16_d64_tail", 2, 2, 2, 77, 100, 64, "NHD", torch.bfloat16),
]


def make_inputs(case, dev):
    name, B, Hq, Hkv, Nq, Nk, D, layout, dt = case
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    shp = lambda h, n: (B, h, n, D) if layout == "HND" else (B, n, h, D)
    q = (torch.randn(shp(Hq, Nq), generator=g) * 1.5).to(dt)
    kb = torch.randn((B, Hkv, 1, D) if layout == "HND" else (B, 1, Hkv, D), generator=g) * 2.0
    k = (torch.randn(shp(Hkv, Nk), generator=g) + kb).to(dt)
    v = (torch.randn(shp(Hkv, Nk), generator=g) + 0.5 * kb).to(dt)
    # one all-zero Q block and one constant K block exercise the 1e-7 amax floor (fused.cu:147-160)
    if layout == "HND":
        q[0, 0, :128] = 0
    else:
        q[0, :128, 0] = 0
    return q.to(dev), k.to(dev), v.to(dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--variant", default="ieee", choices=["ieee", "fast"])
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    dev = torch.device("cuda:0")
    F = BRF.load("ieee")
    Ff = BRF.load("fast") if args.variant == "fast" else None
    report = []
    for case in CASES:
        name, B, Hq, Hkv, Nq, Nk, D, layout, dt = case
        q, k, v = make_inputs(case, dev)
        seq = 2 if layout == "HND" else 1
        # km as the caller hands it over: k.mean(dim=seq, keepdim=True) in k's dtype (core.py:293).  The golden
        # stores it, so the pinned arithmetic starts at the kernel boundary (the mean's last bit is Paddle's).
        km = k.float().mean(dim=seq, keepdim=True).to(dt)
        sm = 1.0 / D ** 0.5
        out = {"q": _bits(q), "k": _bits(k), "v": _bits(v), "km": _bits(km), "sm_scale": np.float64(sm),
               "layout": layout, "dtype": str(dt).split(".")[-1]}
        qi, qs, ki, ks = ref_per_block_int8(F, q, k, km, 128, 64, sm, layout)
        out.update(pb_q_int8=_bits(qi), pb_q_scale=_bits(qs), pb_k_int8=_bits(ki), pb_k_scale=_bits(ks))
        qi2, qs2, ki2, ks2 = ref_per_block_int8(F, q, k, None, 128, 64, sm, layout)
        out.update(pbn_k_int8=_bits(ki2), pbn_k_scale=_bits(ks2))
        qw, qws, kw, kws = ref_per_warp_int8(F, q, k, km, layout)
        out.update(pw_q_int8=_bits(qw), pw_q_scale=_bits(qws), pw_k_int8=_bits(kw), pw_k_scale=_bits(kws))
        vmean = v.float().mean(dim=seq).to(dt)
        vsm = torch.empty(v.shape, dtype=torch.float16, device=dev)
        F.sub_mean_cuda(v, vmean, vsm, 0 if layout == "NHD" else 1)
        out.update(sm_vm=_bits(vmean), sm_v=_bits(vsm))
        for smooth in (False, True):
            vt, v8, vs, vm = ref_per_channel_fp8(F, v, layout, 448.0, smooth)
            tag = "f8s" if smooth else "f8"
            out.update({f"{tag}_v8": _bits(v8), f"{tag}_scale": _bits(vs)})
            if vm is not None:
                out[f"{tag}_vm"] = _bits(vm)
        out["f8_vt"] = _bits(vt)
        torch.cuda.synchronize()
        np.savez_compressed(os.path.join(args.out, f"fused_{name}.npz"), **out)
        line = f"{name}: wrote {len(out)} arrays"
        if Ff is not None:
            a = ref_per_block_int8(Ff, q, k, km, 128, 64, sm, layout)
            b = ref_per_block_int8(F, q, k, km, 128, 64, sm, layout)
            dq = int((a[0] != b[0]).sum()); dk = int((a[2] != b[2]).sum())
            dsq = int((a[1] != b[1]).sum()); dsk = int((a[3] != b[3]).sum())
            _, v8f, vsf, _ = ref_per_channel_fp8(Ff, v, layout, 448.0, False)
            _, v8i, vsi, _ = ref_per_channel_fp8(F, v, layout, 448.0, False)
            d8 = int((v8f.view(torch.uint8) != v8i.view(torch.uint8)).sum())
            line += (f"; --use_fast_math vs IEEE build: q codes {dq}/{a[0].numel()}, k codes {dk}/{a[2].numel()}, "
                     f"q scales {dsq}, k scales {dsk}, fp8 codes {d8}/{v8i.numel()}")
        print(line, flush=True)
        report.append(line)
    with open(os.path.join(args.out, "fused_golden_report.txt"), "w") as f:
        f.write("\n".join(report) + "\n")


if __name__ == "__main__":
    main()
