#!/usr/bin/env python
"""Reference-on-the-same-B200 measurement (SURVEY 8d "also report"): JIT-compiles the reference's own, unmodified
Triton kernels (staged by oracle/stage_ref.sh under oracle/_ref/reference, git-ignored) with Triton on the GPU box and

  * times them (quantizers + attention, CUDA events, 5 warm-ups, 20 reps) next to this repo's kernels,
  * compares this repo's outputs with the reference kernels' outputs AT FULL SIZE (codes / scales bit-exact;
    attention cos-sim + max-abs, the tolerance north_star states: cos-sim >= 0.999),
  * adds flash_attn 2.8 / torch SDPA fp16 as context lines and FP32-SDPA cos-sim on a slice.

Measurement aid only: not imported by tests/, bench.py or smoke().  Output: JSON lines on stdout.
usage: ref_on_b200.py [c2 c2c c1 ...]"""
import json
import os
import sys

os.environ["TRITON_INTERPRET"] = "0"
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
os.environ.setdefault("LOWBIT_REFERENCE_ROOT", os.path.join(ROOT, "oracle", "_ref", "reference"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import lowbit_quant_fa2_paddle_b200 as L  # noqa: E402
from oracle import ref_triton as RT  # noqa: E402

SHAPES = {"c1": (1, 2, 512, 64, False), "c2": (4, 32, 4096, 64, False), "c2c": (4, 32, 4096, 64, True),
          "d128": (4, 32, 4096, 128, False), "d128c8k": (4, 32, 8192, 128, True), "c2_8k": (4, 32, 8192, 64, False)}
dev = torch.device("cuda:0")


def timed(f, reps=20, warm=5):
    for _ in range(warm):
        f()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm()))


def main():
    if not RT.available():
        print(json.dumps({"unavailable": "reference kernels not staged (run oracle/stage_ref.sh in the build container)"}))
        return
    for name in (sys.argv[1:] or ["c1", "c2", "c2c"]):
        b, h, n, d, causal = SHAPES[name]
        torch.manual_seed(0)
        q, k, v = (torch.randn(b, h, n, d, dtype=torch.float16, device=dev) for _ in range(3))
        ops = 4 * b * h * n * n * d / (2 if causal else 1)
        rec = {"shape": name, "B": b, "H": h, "N": n, "D": d, "causal": causal}
        # ---- quantizers: reference Triton (k - km as a Paddle-style elementwise op, then the two kernels) vs ours
        km = L.k_mean(k)
        r_q = RT.per_block_int8(q, k, km=km)
        o_q = L.per_block_int8(q, k, km=km)
        for backend in ("triton", "triton_gpu"):
            g_q = L.per_block_int8(q, k, km=km, backend=backend)
            rec[f"quant[{backend}]_bit_exact_vs_jit_reference(q,qs,k,ks)"] = [bool(torch.equal(a, c)) for a, c in zip(g_q, r_q)]
            rec[f"quant[{backend}]_mismatching_codes(q,k)"] = [int((g_q[i] != r_q[i]).sum()) for i in (0, 2)]
            rec[f"quant[{backend}]_max_code_diff(q,k)"] = [int((g_q[i].int() - r_q[i].int()).abs().max()) for i in (0, 2)]
            rec[f"quant[{backend}]_mismatching_scales(q,k)"] = [int((g_q[i] != r_q[i]).sum()) for i in (1, 3)]
        r4 = (RT.quant_per_block(q, 128, (d ** -0.5) * 1.44269504, "HND", 4), RT.quant_per_block(k - km, 64, 1.0, "HND", 4))
        g4 = L.per_block_int4_unpack(q, k, km=km, backend="triton_gpu")
        rec["quant4[triton_gpu]_bit_exact_vs_jit_reference(q,qs,k,ks)"] = [
            bool(torch.equal(g4[0], r4[0][0])), bool(torch.equal(g4[1], r4[0][1])),
            bool(torch.equal(g4[2], r4[1][0])), bool(torch.equal(g4[3], r4[1][1]))]
        rec["ref_quant_ms"] = timed(lambda: RT.per_block_int8(q, k, km=k.mean(dim=2, keepdim=True)))
        rec["our_quant_ms"] = timed(lambda: L.per_block_int8(q, k, km=L.k_mean(k)))
        # ---- attention kernel only, same codes
        qi, qs, ki, ks = o_q
        r_o, r_lse = RT.attn_forward(qi, ki, v, qs, ks, "HND", causal, torch.float16, True)
        fwd = L.forward_causal if causal else L.forward
        o_o, o_lse = fwd(qi, ki, v, qs, ks, tensor_layout="HND", output_dtype=torch.float16, return_lse=True)
        rec["attn_cos_vs_reference_kernel"] = cos(o_o, r_o)
        rec["attn_max_abs_vs_reference_kernel"] = float((o_o.float() - r_o.float()).abs().max())
        rec["lse2_max_abs_vs_reference_kernel"] = float((o_lse - r_lse).abs().max())
        t_ref = timed(lambda: RT.attn_forward(qi, ki, v, qs, ks, "HND", causal, torch.float16, False))
        t_our = timed(lambda: fwd(qi, ki, v, qs, ks, tensor_layout="HND", output_dtype=torch.float16))
        rec["ref_attn_ms"], rec["ref_attn_tops"] = t_ref, ops / t_ref / 1e9
        rec["our_attn_ms"], rec["our_attn_tops"] = t_our, ops / t_our / 1e9
        # ---- whole operator (quantize + attention)
        api = lambda: L.lowbit_fa_qk_int8_pv_fp16_triton(q, k, v, is_causal=causal)  # noqa: E731

        def ref_api():
            kmm = k.mean(dim=2, keepdim=True)
            a, s1, c, s2 = RT.per_block_int8(q, k, km=kmm)
            return RT.attn_forward(a, c, v, s1, s2, "HND", causal, torch.float16, False)[0]
        t_ref_e, t_our_e = timed(ref_api), timed(api)
        rec["ref_op_ms"], rec["ref_op_tops"] = t_ref_e, ops / t_ref_e / 1e9
        rec["our_op_ms"], rec["our_op_tops"] = t_our_e, ops / t_our_e / 1e9
        o_full, r_full = api(), ref_api()
        rec["op_cos_vs_reference"] = cos(o_full, r_full)
        rec["op_max_abs_vs_reference"] = float((o_full.float() - r_full.float()).abs().max())
        # ---- accuracy vs FP32 SDPA on a slice (1 batch x 4 heads)
        hs = min(h, 4)
        ref32 = torch.nn.functional.scaled_dot_product_attention(q[:1, :hs].float(), k[:1, :hs].float(),
                                                                 v[:1, :hs].float(), is_causal=causal)
        rec["our_cos_vs_fp32_sdpa"] = cos(o_full[:1, :hs], ref32)
        rec["ref_cos_vs_fp32_sdpa"] = cos(r_full[:1, :hs], ref32)
        # ---- context lines
        t_sdpa = timed(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal))
        rec["torch_sdpa_fp16_tops"] = ops / t_sdpa / 1e9
        try:
            from flash_attn import flash_attn_func
            qn, kn, vn = (t.transpose(1, 2).contiguous() for t in (q, k, v))
            t_fa = timed(lambda: flash_attn_func(qn, kn, vn, causal=causal))
            rec["flash_attn2_fp16_tops"] = ops / t_fa / 1e9
        except Exception as e:  # noqa: BLE001
            rec["flash_attn2_fp16_tops"] = f"unavailable: {type(e).__name__}"
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
