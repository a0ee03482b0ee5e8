#!/usr/bin/env python
"""Per-phase cycle counts of the softmax warps of attn_fwd_n64_kernel at config 2 (development aid).
Needs a trace build:  make -C lowbit_quant_fa2_paddle_b200/csrc clean all EXTRA=-DLOWBIT_TRACE  (the DBG instantiation
then writes %clock stamps of one CTA -- block (16, 5, 1), lane 0 of each softmax warp -- into the debug buffer instead of
the integer scores); rebuild without EXTRA afterwards.  Output of the last run: profiles/r2_attn_c2_phase_trace.txt."""
import os, sys, torch, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import lowbit_quant_fa2_paddle_b200 as L
from lowbit_quant_fa2_paddle_b200 import attention as A, _native as NV
dev = torch.device("cuda:0")
torch.manual_seed(0)
q, k, v = (torch.randn(4, 32, 4096, 64, dtype=torch.float16, device=dev) for _ in range(3))
km = L.k_mean(k)
qi, qs, ki, ks = L.per_block_int8(q, k, km=km)
f = lambda: A._forward(qi, ki, v, qs, ks, "HND", torch.float16, False, False)
for _ in range(3): f()
buf = torch.zeros(128 * 64, dtype=torch.int32, device=dev)
lib = NV.lib()
lib.lowbit_attn_set_debug_buffer.argtypes = [ctypes.c_void_p]
lib.lowbit_attn_set_debug_buffer(ctypes.c_void_p(buf.data_ptr()))
f(); torch.cuda.synchronize()
lib.lowbit_attn_set_debug_buffer(ctypes.c_void_p(0))
t = buf.cpu().view(-1, 64, 8)[:4].to(torch.int64)   # [warp][block][stamp]
names = ["P_lo chunk", "wait S_hi", "ld S_hi", "P_hi chunk", "st P + wait S_lo(j+1)", "ld S_lo(j+1)", "wait st + arrive"]
for w in range(4):
    d = (t[w, :, 1:] - t[w, :, :-1]) & 0xffffffff
    blk = (t[w, 1:, 0] - t[w, :-1, 0]) & 0xffffffff
    print(f"warp {w}: cycles per block median {blk[4:60].median().item()} mean {blk[4:60].float().mean().item():.0f}")
    for i, n in enumerate(names):
        print(f"   {n:24s} median {d[4:60, i].median().item():6d}  mean {d[4:60, i].float().mean().item():7.0f}")
    gap = (t[w, 1:, 0] - t[w, :-1, 7]) & 0xffffffff
    print(f"   {'loop overhead':24s} median {gap[4:60].median().item():6d}")
