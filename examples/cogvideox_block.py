#!/usr/bin/env python
"""Synthetic CogVideoX-5B attention block with the low-bit operator patched in as scaled_dot_product_attention.

The reference's end-to-end claim (README.md:24, example/sageattn_cogvideo.py) needs the CogVideoX weights, which this
environment cannot fetch; the plug-in PATH does not: this script builds one transformer attention block of the 5B
model's geometry (hidden 3072 = 48 heads x 64, 226 text + 17550 video tokens = 17776, bf16, random weights), runs it
once with torch's SDPA and once under `patch_sdpa`, and reports the block-output error and both timings.  The linear
layers and the layer norms are ordinary torch modules (library kernels): they are the model, not the hot path.

    python examples/cogvideox_block.py [--op int8|q8k4|int4] [--batch 2] [--tokens 17776]
"""
import argparse
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import lowbit_quant_fa2_paddle_b200 as L  # noqa: E402
from lowbit_quant_fa2_paddle_b200.plugin import patch_sdpa  # noqa: E402


class Attention(nn.Module):
    """CogVideoX-style joint text/video self-attention: LayerNorm -> qkv -> per-head q/k LayerNorm -> SDPA -> out."""

    def __init__(self, dim=3072, heads=48):
        super().__init__()
        self.heads, self.hd = heads, dim // heads
        self.norm = nn.LayerNorm(dim)
        self.to_qkv = nn.Linear(dim, 3 * dim)
        self.norm_q, self.norm_k = nn.LayerNorm(self.hd), nn.LayerNorm(self.hd)
        self.to_out = nn.Linear(dim, dim)

    def forward(self, x):
        b, n, _ = x.shape
        qkv = self.to_qkv(self.norm(x)).view(b, n, 3, self.heads, self.hd).permute(2, 0, 3, 1, 4)  # [3, B, H, N, D]
        q, k, v = self.norm_q(qkv[0]), self.norm_k(qkv[1]), qkv[2]
        # diffusers calls SDPA positionally with attn_mask / dropout_p / is_causal keywords
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
        return x + self.to_out(o.transpose(1, 2).reshape(b, n, -1))


def timed(f, reps=5):
    for _ in range(2):
        f()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        y = f()
    b.record()
    torch.cuda.synchronize()
    return y, a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--op", default="int8", choices=["int8", "q8k4", "int4"])
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--tokens", type=int, default=17776)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    blk = Attention().to(dev, torch.bfloat16).eval()
    x = torch.randn(args.batch, args.tokens, 3072, device=dev, dtype=torch.bfloat16)
    op = {"int8": L.lowbit_fa_qk_int8_pv_fp16_triton, "q8k4": L.lowbit_fa_q_int8_k_int4_pv_fp16,
          "int4": L.lowbit_fa_qk_int4_pv_fp16_triton}[args.op]
    with torch.no_grad():
        y_ref, t_ref = timed(lambda: blk(x))
        with patch_sdpa(op):
            y_low, t_low = timed(lambda: blk(x))
    d = (y_low.float() - y_ref.float())
    cos = F.cosine_similarity(y_low.float().flatten(), y_ref.float().flatten(), dim=0).item()
    print(f"CogVideoX-5B attention block, B={args.batch} N={args.tokens} bf16, op={args.op}: "
          f"torch SDPA {t_ref:.2f} ms, low-bit {t_low:.2f} ms ({t_ref / t_low:.2f}x); block output vs torch SDPA: "
          f"cos {cos:.6f}, MSE {d.pow(2).mean().item():.3e}, max |diff| {d.abs().max().item():.3e}")


if __name__ == "__main__":
    main()
