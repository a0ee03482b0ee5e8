#!/usr/bin/env python
"""tools/make_golden_qattn.py -- golden vectors for row A3 (and a CUDA-side cross-check of A1 / A2) from the
reference's OWN CUDA attention kernels.

Runs on a B200.  Imports the extension that oracle/build_ref_qattn.py compiled from
/root/reference/csrc/qattn/{qk_int_sv_f8_cuda.cu, qk_int_sv_f16_cuda.cu} (unmodified; mma.sync code that also runs on
sm_100) and calls it the way src/core.py:882-916 does:

    qk_int8_sv_f8_accum_f32_fuse_v_scale_attn(q_int8, k_int8, v_fp8, o, q_scale, k_scale, v_scale,
                                              tensor_layout, is_causal, qk_quant_gran=2 (per warp), sm_scale, return_lse)
    qk_int8_sv_f16_accum_f32_attn(q_int8, k_int8, v, o, q_scale, k_scale, ...)            (fp16 inputs only)

Inputs are the codes of the committed tests/golden/fused_*.npz fixtures (made by the reference's own quantizers,
tools/make_golden_fused.py): per-block Q codes (one scale per 128 rows, sm_scale * log2e folded in, src/quant.py:70-98),
K codes smoothed by km (one scale per 64 rows), e4m3 V^T with per-channel scales (src/quant.py:254-291).  The kernel
wants one Q scale per 32-row warp and folds sm_scale * log2e in itself, so it is handed every block scale four times
and sm_scale = 1 / log2e (the product with log2e is 1 within one fp32 ulp).  Causal needs qo_len == kv_len, which no
fused fixture has: two more cases are quantized here with the reference's fused.cu kernels and stored whole.

Writes tests/golden/qattn_*.npz (or --out DIR; on the GPU box: gpurun_out/golden_qattn, copied into tests/golden/).
"""
import argparse
import glob
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import build_ref_fused as BRF  # noqa: E402
from oracle import build_ref_qattn as BRQ  # noqa: E402
import make_golden_fused as MGF  # noqa: E402

INV_LOG2E = 1.0 / 1.44269504


def t16(g, name, dt):
    return torch.from_numpy(g[name]).view(dt)


def run_ref(A, layout, causal, qi, qs, ki, ks, v16, v8, vs, dt):
    """-> dict of outputs of the reference kernels on these codes."""
    tl = 0 if layout == "NHD" else 1
    dev = qi.device
    qs_w = qs.repeat_interleave(4, dim=2).contiguous()  # one scale per 32-row warp (CTA_Q 128 / WARP_Q 32)
    out = {}
    o = torch.empty(qi.shape, dtype=dt, device=dev)
    lse = A.qk_int8_sv_f8_accum_f32_fuse_v_scale_attn(qi, ki, v8, o, qs_w, ks, vs, tl, int(causal), 2, INV_LOG2E, 1)
    torch.cuda.synchronize()
    out["o_f8"], out["lse_f8"] = MGF._bits(o), lse.float().cpu().numpy()
    if dt == torch.float16:
        o2 = torch.empty(qi.shape, dtype=dt, device=dev)
        lse2 = A.qk_int8_sv_f16_accum_f32_attn(qi, ki, v16, o2, qs_w, ks, tl, int(causal), 2, INV_LOG2E, 1)
        torch.cuda.synchronize()
        out["o_f16"], out["lse_f16"] = MGF._bits(o2), lse2.float().cpu().numpy()
    return out


CAUSAL_CASES = [
    # name, B, Hq, Hkv, N, D, layout, dtype
    ("causal_hnd_f16_d64", 1, 2, 2, 384, 64, "HND", torch.float16),
    ("causal_nhd_bf16_d128_gqa", 1, 4, 2, 320, 128, "NHD", torch.bfloat16),
    ("causal_hnd_f16_d128_tail", 1, 2, 1, 333, 128, "HND", torch.float16),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    dev = torch.device("cuda:0")
    A = BRQ.load()
    F = BRF.load("ieee")
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "fused_*.npz"))):
        name = os.path.basename(path)[len("fused_"):-len(".npz")]
        g = np.load(path)
        layout, dt = str(g["layout"]), getattr(torch, str(g["dtype"]))
        qi, ki = (torch.from_numpy(g[n]).to(dev) for n in ("pb_q_int8", "pb_k_int8"))
        qs, ks, vs = (torch.from_numpy(g[n]).to(dev) for n in ("pb_q_scale", "pb_k_scale", "f8_scale"))
        v16 = t16(g, "v", dt).to(dev)
        v8 = torch.from_numpy(g["f8_v8"]).view(torch.float8_e4m3fn).to(dev)
        out = run_ref(A, layout, False, qi, qs, ki, ks, v16, v8, vs, dt)
        out.update(inputs=f"fused_{name}.npz", causal=False)
        np.savez_compressed(os.path.join(args.out, f"qattn_{name}.npz"), **out)
        print(f"{name}: {sorted(out)}", flush=True)
    for name, B, Hq, Hkv, N, D, layout, dt in CAUSAL_CASES:
        q, k, v = MGF.make_inputs((name, B, Hq, Hkv, N, N, D, layout, dt), dev)
        seq = 2 if layout == "HND" else 1
        km = k.float().mean(dim=seq, keepdim=True).to(dt)
        sm = 1.0 / D ** 0.5
        qi, qs, ki, ks = MGF.ref_per_block_int8(F, q, k, km, 128, 64, sm, layout)
        _, v8, vs, _ = MGF.ref_per_channel_fp8(F, v, layout, 448.0, False)
        out = run_ref(A, layout, True, qi, qs, ki, ks, v, v8, vs, dt)
        out.update(causal=True, layout=layout, dtype=str(dt).split(".")[-1], q=MGF._bits(q), k=MGF._bits(k), v=MGF._bits(v),
                   km=MGF._bits(km), sm_scale=np.float64(sm), pb_q_int8=MGF._bits(qi), pb_q_scale=MGF._bits(qs),
                   pb_k_int8=MGF._bits(ki), pb_k_scale=MGF._bits(ks), f8_v8=MGF._bits(v8), f8_scale=MGF._bits(vs))
        np.savez_compressed(os.path.join(args.out, f"qattn_{name}.npz"), **out)
        print(f"{name}: {sorted(out)}", flush=True)


if __name__ == "__main__":
    main()
