"""CPU (gloo, world_size 2 and 4): the host side of the multi-GPU partitionings -- head sharding arithmetic, the
ring schedule (contiguous and zig-zag chunks, causal skipping), the flat K/V message layout, the double-buffered
P2P exchange and the state merge order -- with the CUDA kernels replaced by the CPU oracle through the backend
hook of parallel.ring_attention.  The product path never takes this hook (parallel.CudaBackend is the default and
refuses CPU tensors)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import cos_sim

from lowbit_quant_fa2_paddle_b200 import parallel as P


class OracleBackend:
    """Same interface as parallel.CudaBackend, oracle math on CPU tensors (test infrastructure)."""

    def __init__(self, layout, qk, pv, sm_scale):
        from oracle import attention as OA
        from oracle import quant as OQ
        self.OA, self.OQ = OA, OQ
        self.layout, self.qk, self.pv, self.sm_scale = layout, qk, pv, sm_scale

    def k_sum(self, k, seq_dim):
        return k.sum(dim=seq_dim, dtype=torch.float64)

    def quantize_q(self, q_chunk):
        return self.OQ.quant_per_block_q1(q_chunk, 128, self.sm_scale * 1.44269504, self.layout, bits=8)

    def quantize_kv(self, k_chunk, v_chunk, km, msg, i):
        seq = 2 if self.layout == "HND" else 1
        ks = k_chunk if km is None else (k_chunk - km.unsqueeze(seq)).to(k_chunk.dtype)  # `k - km` in the input dtype
        if self.qk == "mixed":
            codes, scale, kb = self.OQ.quant_k_mixed(ks, None, None, 64, self.layout)
            msg.view(f"k{i}").copy_(self.OQ.pack_mixed(codes, kb, 64, self.layout))
            msg.view(f"kb{i}").copy_(kb)
        else:
            bits = 4 if self.qk == "int4" else 8
            codes, scale = self.OQ.quant_per_block_q1(ks, 64, 1.0, self.layout, bits=bits)
            msg.view(f"k{i}").copy_(self.OQ.pack_codes(codes, 4) if bits == 4 else codes)
        msg.view(f"ks{i}").copy_(scale)
        if self.pv == "fp8":
            v8, vs, _ = self.OQ.per_channel_fp8(v_chunk, self.layout, smooth_v=False)
            msg.view(f"v{i}").copy_(v8)
            msg.view(f"vs{i}").copy_(vs)
        else:
            msg.view(f"v{i}").copy_(v_chunk)

    def partial(self, state, q_pack, msg, i, q_off, k_off, causal):
        kc = msg.view(f"k{i}")
        if self.qk == "int4":
            kc = self.OQ.unpack_codes(kc, 4)
        elif self.qk == "mixed":
            kc = self.OQ.unpack_mixed(kc, msg.view(f"kb{i}"), 64, self.layout)
        v = msg.view(f"v{i}")
        vs = None
        if self.pv == "fp8":
            n = kc.shape[2] if self.layout == "HND" else kc.shape[1]
            v, vs = self.OA.v8_to_natural(v, n, self.layout), msg.view(f"vs{i}")
        return self.OA.attn_partial(state, q_pack[0], kc, v, q_pack[1], msg.view(f"ks{i}"), self.layout, causal,
                                    q_off, k_off, v_scale=vs)

    def finalize(self, state, q_pack, out_dtype, return_lse):
        o, lse2 = self.OA.attn_finalize(state, q_pack[0], self.layout, out_dtype)
        return o, (lse2 if return_lse else None)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make_qkv(cfg):
    layout, causal, qk, pv, zigzag, d, hq, hkv, n = cfg
    torch.manual_seed(1234)
    seq = 2 if layout == "HND" else 1
    shp = lambda h: (1, h, n, d) if layout == "HND" else (1, n, h, d)
    q = torch.randn(shp(hq)).half()
    k = (torch.randn(shp(hkv)) + 2.0 * torch.randn(shp(hkv)[:seq] + (1,) + shp(hkv)[seq + 1:])).half()
    v = torch.randn(shp(hkv)).half()
    if qk == "mixed":  # block magnitudes spread over the three width classes (around the common mean)
        km = k.float().mean(dim=seq, keepdim=True)
        w = torch.ones(n)
        w[n // 4: n // 2] = 0.15
        w[n // 2: 3 * n // 4] = 3.0
        w = w.view(1, 1, n, 1) if layout == "HND" else w.view(1, n, 1, 1)
        k = (km + (k.float() - km) * w).half()
    return q, k, v


def _worker(rank, world, port, cfg, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        layout, causal, qk, pv, zigzag, d, hq, hkv, n = cfg
        seq = 2 if layout == "HND" else 1
        q, k, v = _make_qkv(cfg)  # every rank builds the same full tensors, then keeps its shard
        chunks = P.seq_chunks(n, world, rank, zigzag)
        take = lambda x: torch.cat([x.narrow(seq, c.offset, c.length) for c in chunks], dim=seq).contiguous()
        be = OracleBackend(layout, qk, pv, d ** -0.5)
        # the NHD configurations run over a ring_group (own process group with a short timeout) in check mode
        grp = P.ring_group(timeout_s=120.0, backend="gloo") if layout == "NHD" else None
        o, lse = P.ring_attention(take(q), take(k), take(v), tensor_layout=layout, is_causal=causal, qk=qk, pv=pv,
                                  zigzag=zigzag, return_lse=True, backend=be, group=grp, check=grp is not None)
        ret[rank] = (o, lse, [(c.offset, c.length) for c in chunks])
    finally:
        dist.destroy_process_group()


RING_CFGS = [
    # layout, causal, qk, pv, zigzag, d, hq, hkv, n
    ("HND", False, "int8", "fp16", False, 64, 2, 2, 512),
    ("NHD", True, "int4", "fp16", True, 64, 4, 2, 512),
    ("HND", True, "int4", "fp8", True, 128, 2, 1, 512),
    ("NHD", True, "int8", "fp16", False, 64, 2, 2, 256),
    ("HND", True, "mixed", "fp16", True, 128, 2, 2, 512),   # dynamic INT8/INT4/INT2 K blocks in the D-byte container
]


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("cfg", RING_CFGS)
def test_ring_attention_gloo_matches_single_process_oracle(cfg, world):
    from oracle import attention as OA
    layout, causal, qk, pv, zigzag, d, hq, hkv, n = cfg
    if n // (world * (2 if zigzag else 1)) % 128 != 0:
        # chunks must be whole 128-row Q blocks for the per-chunk Q scales to equal the single-process ones
        cfg = cfg[:-1] + (128 * world * (2 if zigzag else 1),)
        n = cfg[-1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), cfg, ret), nprocs=world, join=True)
    # reassemble the ring output in global token order
    seq = 2 if layout == "HND" else 1
    q, k, v = _make_qkv(cfg)
    o_ring = torch.empty_like(q)
    lse_ring = torch.empty(1, hq, n)
    for r in range(world):
        o, lse, chunks = ret[r]
        pos = 0
        for off, ln in chunks:
            o_ring.narrow(seq, off, ln).copy_(o.narrow(seq, pos, ln))
            lse_ring[:, :, off:off + ln] = lse[:, :, pos:pos + ln]
            pos += ln
    o_ref, lse_ref = OA.lowbit_fa_api(q, k, v, layout, causal, return_lse=True, compat_tail=False, pv_accum="fp32",
                                      qk=qk, pv=pv)
    diff = (o_ring.float() - o_ref.float()).abs()
    if pv == "fp16":
        assert diff.max() <= 4e-3
    else:  # fp8: per-shard v_scale -> V codes differ by up to one e4m3 step (2^-3 of the value) from the global ones
        assert diff.max() <= 0.125 * float(v.float().abs().max())
    assert cos_sim(o_ring, o_ref) >= (0.9999 if pv == "fp16" else 0.998)
    assert (lse_ring - lse_ref).abs().max() <= (1e-2 if pv == "fp16" else 5e-2)


def test_head_shard_arithmetic():
    assert P.head_shard(48, 48, 8, 3) == (18, 24, 18, 24)
    assert P.head_shard(32, 8, 4, 1) == (8, 16, 2, 4)   # GQA: a group stays with its kv head
    with pytest.raises(ValueError):
        P.head_shard(48, 48, 5, 0)
    seen = []
    for r in range(8):
        hq0, hq1, _, _ = P.head_shard(48, 48, 8, r)
        seen += list(range(hq0, hq1))
    assert seen == list(range(48))


def test_ring_schedule_covers_every_pair_once_and_skips_the_future():
    for world in (2, 4, 8):
        for zigzag in (False, True):
            n = 64 * 2 * world * 2
            for causal in (False, True):
                work = []
                for rank in range(world):
                    mine = P.seq_chunks(n, world, rank, zigzag)
                    cnt = 0
                    seen_src = set()
                    for step in range(world):
                        src = P.ring_source(rank, step, world)
                        seen_src.add(src)
                        for qc in mine:
                            for kc in P.seq_chunks(n, world, src, zigzag):
                                if P.pair_visible(qc, kc, causal):
                                    # visible pairs really contain a visible (row, key) pair; skipped ones do not
                                    assert (not causal) or kc.offset <= qc.offset + qc.length - 1
                                    lo = min(qc.offset + qc.length, kc.offset + kc.length)
                                    cnt += qc.length * kc.length if not causal else max(0, lo - kc.offset) * qc.length
                                else:
                                    assert kc.offset > qc.offset + qc.length - 1
                    assert seen_src == set(range(world))
                    work.append(cnt)
                tokens = sorted(t for r in range(world) for c in P.seq_chunks(n, world, r, zigzag)
                                for t in range(c.offset, c.offset + c.length))
                assert tokens == list(range(n))  # the chunks partition the sequence
                if causal and zigzag:
                    assert max(work) == min(work)  # zig-zag balances causal work exactly at chunk granularity


def test_ring_message_layout_roundtrip():
    fs = P.kv_fields(2, 3, 64, [P.Chunk(0, 128), P.Chunk(896, 128)], "NHD", "int4", "fp8")
    msg = P.RingMessage(fs, torch.device("cpu"))
    names = [f.name for f in fs]
    assert names == ["k0", "ks0", "v0", "vs0", "k1", "ks1", "v1", "vs1"]
    assert msg.view("k0").shape == (2, 128, 3, 32) and msg.view("v1").shape == (2, 64, 3, 128)
    for f in fs:
        assert f.offset % 256 == 0
    # views alias the flat buffer and do not overlap
    msg.flat.zero_()
    msg.view("ks1").fill_(1.5)
    assert float(msg.view("ks1").sum()) == 1.5 * 2 * 3 * 2 and int((msg.flat != 0).sum()) == 2 * 3 * 2 * 4 - (2 * 3 * 2 * 4) // 4 * 2
    other = msg.like()
    assert other.flat.data_ptr() != msg.flat.data_ptr() and other.view("k1").shape == msg.view("k1").shape


def test_ring_refuses_cpu_tensors_without_a_test_backend():
    """No CPU fallback: the product path (default backend) raises on CPU tensors before touching any collective."""
    from lowbit_quant_fa2_paddle_b200 import _native
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        x = torch.randn(1, 2, 128, 64).half()
        with pytest.raises(_native.LowbitNativeError):
            P.ring_attention(x, x, x)
    finally:
        dist.destroy_process_group()
