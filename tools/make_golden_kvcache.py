#!/usr/bin/env python
"""Golden vectors for the KV-cache attention path, made by the reference's own code.

The prototype kernel `_fwd_kernel` (src/triton/quantization/attn_4bit_per_block.py) does not run as written, but its
driver (`main`, :637-788) states what it is checked against: FlashAttention over the caches dequantized by the
reference's `unpack_and_dequant_kcache` / `unpack_and_dequant_vcache` (src/triton/utils/quant/new_pack.py:68-144) --
`err_o = (out1 - out2).abs().mean()`.  This script reproduces that check path with the reference's code UNMODIFIED:

  K, V (fp16)  --  triton_quantize_and_pack_along_last_dim (its two Triton kernels under TRITON_INTERPRET=1, exactly as
                   tools/make_golden.py does for the kivi_* fixtures; K transposed to [B,D,H,N] like the driver :655-661)
               --  unpack_and_dequant_kcache / unpack_and_dequant_vcache, imported from the reference file and executed
                   on torch tensors through a Paddle shim (the functions only use zeros / arange / view / unsqueeze /
                   shifts / casts; the shim supplies `Tensor.place` and Paddle's list-of-slices indexing)
               --  exact softmax attention (fp64) over the dequantized fp16 tensors = what the driver's FlashAttention
                   call approximates.

Run in the build container (needs /root/reference):  python tools/make_golden_kvcache.py  ->  tests/golden/kvcache_*.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import ref_triton as R  # noqa: E402


def paddle_shim():
    """What the two dequantizers need from `paddle`, on torch tensors."""
    import paddle  # the stub module oracle.ref_triton installed
    paddle.int16, paddle.int32 = torch.int16, torch.int32
    paddle.zeros = lambda shape, dtype=None, device=None: torch.zeros(tuple(shape), dtype=dtype, device=device)
    paddle.arange = lambda n, device=None: torch.arange(n, device=device)
    torch.Tensor.place = property(lambda self: self.device)
    orig = torch.Tensor.__getitem__

    def getitem(self, idx):  # Paddle accepts a LIST of slices / index tensors as a multi-dimensional index
        return orig(self, tuple(idx) if isinstance(idx, list) else idx)
    torch.Tensor.__getitem__ = getitem


def main():
    npk = R._load("src/triton/utils/quant/new_pack.py")   # the reference module itself
    paddle_shim()
    out_dir = os.path.join(ROOT, "tests", "golden")
    cases = [("d128_b4", 1, 2, 2, 256, 128, 4), ("d64_b4_q3", 1, 3, 2, 192, 64, 4), ("d128_b2", 1, 1, 2, 128, 128, 2)]
    for name, B, Nq, H, N, D, bits in cases:
        g = torch.Generator().manual_seed(4200 + D + bits)
        q = torch.randn(B, Nq, H, D, generator=g).to(torch.float16)
        k = (torch.randn(B, N, H, D, generator=g) + 0.5 * torch.randn(1, 1, H, D, generator=g)).to(torch.float16)
        v = torch.randn(B, N, H, D, generator=g).to(torch.float16)
        kt = k.permute(0, 3, 2, 1).contiguous()                      # [B,D,H,N]: `key_states.transpose(1, 3)`
        kcode, kscale, kmn = R.kivi_quantize_and_pack(kt, 32, bits)
        vcode, vscale, vmn = R.kivi_quantize_and_pack(v.contiguous(), 32, bits)
        dk = npk.unpack_and_dequant_kcache(kcode, kscale, kmn, group_size=32, bits=bits)   # [B,D,H,N] fp16
        dv = npk.unpack_and_dequant_vcache(vcode, vscale, vmn, group_size=32, bits=bits)   # [B,N,H,D] fp16
        assert dk.dtype == torch.float16 and dv.dtype == torch.float16 and dk.shape == kt.shape and dv.shape == v.shape
        sm = D ** -0.5
        s = torch.einsum("bqhd,bdhn->bhqn", q.double(), dk.double()) * sm
        m = s.amax(dim=-1, keepdim=True)
        p = torch.exp(s - m)
        l = p.sum(dim=-1, keepdim=True)
        o = torch.einsum("bhqn,bnhd->bqhd", p / l, dv.double())
        lse = (m + torch.log(l)).squeeze(-1)
        f16 = lambda t: t.contiguous().view(torch.int16).numpy().view(np.uint16)
        np.savez_compressed(os.path.join(out_dir, f"kvcache_{name}.npz"),
                            q__float16=f16(q), kcode=kcode.numpy(), kscale__float16=f16(kscale), kmn__float16=f16(kmn),
                            vcode=vcode.numpy(), vscale__float16=f16(vscale), vmn__float16=f16(vmn),
                            dequant_k__float16=f16(dk), dequant_v__float16=f16(dv),
                            o=o.float().numpy(), lse=lse.float().numpy(), bits=bits, group_size=32, sm_scale=sm)
        err_k = (kt.float() - dk.float()).abs().mean().item()
        print(f"kvcache_{name}: K quant error {err_k:.4f}, o range {o.abs().max().item():.3f}", flush=True)


if __name__ == "__main__":
    main()
