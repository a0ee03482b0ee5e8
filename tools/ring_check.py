#!/usr/bin/env python
"""Multi-GPU check of parallel.ring_attention and head sharding against the single-GPU kernel.
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/ring_check.py
Every rank builds the same full tensors (seeded), keeps its shard, runs the ring, and rank 0 compares the gathered
result with the single-GPU call on the full tensors.  Prints one line per config and exits non-zero on mismatch."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import lowbit_quant_fa2_paddle_b200 as L  # noqa: E402
from lowbit_quant_fa2_paddle_b200 import parallel as P  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    cfgs = [
        # layout, causal, qk, pv, d, b, hq, hkv, n
        ("HND", False, "int8", "fp16", 64, 1, 4, 4, 2048),
        ("NHD", True, "int4", "fp16", 64, 2, 4, 2, 4096),
        ("HND", True, "int4", "fp8", 128, 1, 4, 4, 4096),
        ("NHD", False, "int8", "fp8", 64, 1, 4, 2, 4096),
        ("HND", True, "int4", "fp16", 128, 1, 8, 8, 16384),
        ("HND", True, "mixed", "fp16", 128, 1, 4, 4, 8192),   # dynamic INT8/INT4/INT2 K blocks
    ]
    for layout, causal, qk, pv, d, b, hq, hkv, n in cfgs:
        torch.manual_seed(7)
        seq = 2 if layout == "HND" else 1
        shp = lambda h: (b, h, n, d) if layout == "HND" else (b, n, h, d)
        q = torch.randn(shp(hq), dtype=torch.float16, device=dev)
        k = torch.randn(shp(hkv), dtype=torch.float16, device=dev) + 1.5
        v = torch.randn(shp(hkv), dtype=torch.float16, device=dev)
        if qk == "mixed":  # block magnitudes over the three width classes (around the common mean)
            w = torch.tensor([0.1, 1.0, 3.0], device=dev)[torch.arange(n, device=dev) // 64 % 3]
            w = w.view(1, 1, n, 1) if layout == "HND" else w.view(1, n, 1, 1)
            k = ((k.float() - 1.5) * w + 1.5).half()
        zig = causal
        chunks = P.seq_chunks(n, world, rank, zig)
        take = lambda x: torch.cat([x.narrow(seq, c.offset, c.length) for c in chunks], dim=seq).contiguous()
        # the NHD configurations run over a ring_group (short timeout) in check mode (per-step error surfacing)
        grp = P.ring_group(timeout_s=120.0) if layout == "NHD" else None
        o, lse = P.ring_attention(take(q), take(k), take(v), tensor_layout=layout, is_causal=causal, qk=qk, pv=pv,
                                  return_lse=True, group=grp, check=grp is not None)
        # reference: the single-GPU entry point on the full tensors
        if qk == "mixed":
            fn = L.lowbit_fa_q_int8_k_dynamic
        elif pv == "fp8":
            fn = L.lowbit_fa_qk_int4_pv_fp8 if qk == "int4" else L.lowbit_fa_qk_int8_pv_fp8_cuda
        else:
            fn = L.lowbit_fa_qk_int4_pv_fp16_triton if qk == "int4" else L.lowbit_fa_qk_int8_pv_fp16_triton
        o_ref, lse_ref = fn(q, k, v, tensor_layout=layout, is_causal=causal, return_lse=True)
        o_loc, lse_loc = take(o_ref), torch.cat([lse_ref[:, :, c.offset:c.offset + c.length] for c in chunks], dim=2)
        err = (o.float() - o_loc.float()).abs().max()
        lerr = (lse - lse_loc).abs().max()
        # accuracy against exact attention (fp32 SDPA on the full tensors), ring and single pass side by side
        hd = 1 if layout == "HND" else 2
        to_h = (lambda x: x) if layout == "HND" else (lambda x: x.permute(0, 2, 1, 3))
        kf, vf = to_h(k).float(), to_h(v).float()
        if hq != hkv:
            kf, vf = kf.repeat_interleave(hq // hkv, dim=1), vf.repeat_interleave(hq // hkv, dim=1)
        sd = torch.nn.functional.scaled_dot_product_attention(to_h(q).float(), kf, vf, is_causal=causal)
        sd = sd if layout == "HND" else sd.permute(0, 2, 1, 3)
        sd_loc = take(sd.contiguous())
        e_ring = (o.float() - sd_loc).abs()
        e_one = (o_loc.float() - sd_loc).abs()
        of, o1 = o.float(), o_loc.float()
        dots = torch.stack([(of * sd_loc).sum(), (of ** 2).sum(), (sd_loc ** 2).sum(), (o1 * sd_loc).sum(), (o1 ** 2).sum()]).double()
        dist.all_reduce(dots)
        cos = float(dots[0] / (dots[1].sqrt() * dots[2].sqrt()))
        cos1 = float(dots[3] / (dots[4].sqrt() * dots[2].sqrt()))
        # rows that see few keys (the head of a causal sequence) carry un-averaged e4m3 rounding of P (2^-4 relative per
        # element): the ring / single-pass difference is reported per band of visible keys
        pos = torch.cat([torch.arange(c.offset, c.offset + c.length, device=dev) for c in chunks])
        vis = pos if causal else torch.full_like(pos, n)
        diff = (of - o1).abs()
        dmax = diff.amax(dim=(0, 1, 3)) if layout == "HND" else diff.amax(dim=(0, 2, 3))  # per local row
        zero = torch.zeros((), device=dev)
        band = lambda lo, hi: torch.where((vis >= lo) & (vis < hi), dmax, zero).max()
        stats = torch.stack([err, lerr, e_ring.max(), e_one.max(), band(256, 1 << 30), e_ring.mean(), e_one.mean(),
                             band(0, 64), band(64, 256)])
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        cos_floor = 0.999 if qk == "int8" else 0.98  # INT4 / mixed K is a coarse code by design (single pass: ~0.99)
        if pv == "fp16":
            good = (bool(stats[0] <= 4e-3) and bool(stats[1] <= 1e-2) and cos >= cos_floor and cos >= cos1 - 1e-5)
        else:
            # FP8 P.V: the ring rounds P~ to e4m3 against per-shard reference maxima, a single pass against one running
            # maximum -- two different, equally valid roundings of the same softmax, 2^-4 relative per element, which
            # average out over the keys a row sees.  The bar is therefore accuracy against exact attention: the ring no
            # worse than the single pass (cos-sim within 2e-4, max error within 1.25x, mean error within 1.05x), and
            # the two within 0.05 of each other on every row that sees at least 256 keys.
            good = (cos >= cos_floor and cos >= cos1 - 2e-4 and bool(stats[2] <= 1.25 * stats[3] + 1e-3)
                    and bool(stats[5] <= 1.05 * stats[6] + 1e-4) and bool(stats[4] <= 0.05) and bool(stats[1] <= 5e-2))
        ok = ok and good
        if rank == 0:
            print(f"ring world={world} {layout} causal={causal} qk={qk} pv={pv} d={d} n={n}: max|o-o1|={float(stats[0]):.3e} "
                  f"(rows seeing <64 / 64-255 / >=256 keys: {float(stats[7]):.3e} / {float(stats[8]):.3e} / {float(stats[4]):.3e}) "
                  f"max|lse-lse1|={float(stats[1]):.3e} | vs fp32 SDPA: cos ring {cos:.6f} / single {cos1:.6f}, "
                  f"max err ring {float(stats[2]):.3e} / single {float(stats[3]):.3e}, mean err ring {float(stats[5]):.3e} / "
                  f"single {float(stats[6]):.3e} {'OK' if good else 'MISMATCH'}", flush=True)
    # head sharding: each rank computes its head slice; gathered result == single-GPU result, bit for bit
    torch.manual_seed(9)
    b, n, h, d = 2, 1200, 8 * world, 64
    q, k, v = (torch.randn(b, n, h, d, dtype=torch.float16, device=dev) for _ in range(3))
    o_full = L.lowbit_fa_q_int8_k_int4_pv_fp16(q, k, v, tensor_layout="NHD")
    o_mine = P.lowbit_fa_head_sharded(q, k, v, L.lowbit_fa_q_int8_k_int4_pv_fp16, world, rank, tensor_layout="NHD")
    hq0, hq1, _, _ = P.head_shard(h, h, world, rank)
    same = torch.tensor([int(torch.equal(o_mine, o_full[:, :, hq0:hq1]))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"head-sharded world={world}: slices bit-identical to the single-GPU result: {bool(same.item())}", flush=True)
    ok = ok and bool(same.item())
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
