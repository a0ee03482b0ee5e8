"""Multi-GPU partitioning of the low-bit attention path: one process per GPU (torch.distributed, NCCL over NVLink).

The reference has no distributed code: its only hooks are `return_lse` ("for Ring Attention",
src/core.py:123-125) and the xDiT demo that shards heads (example/parallel_sageattn_cogvideo.py:46-53).  Two
partitionings are built here on the same kernels:

1. head sharding (`head_shard`, `lowbit_fa_head_sharded`): (batch x kv-head) units are independent
   (grid axes 1,2 of every reference kernel, attn_qk_int8_per_block.py:100-102), so each rank takes a slice of
   the kv heads and the q heads of their groups -- NO collective on the data path.
2. sequence-parallel ring (`ring_attention`): Q, K, V sharded along the sequence.  One tiny all-reduce makes the
   K mean global (smoothing needs the mean over ALL keys); each rank then quantizes its K (INT8 / packed INT4)
   and V (fp16 / e4m3) shard ONCE into a flat message buffer; the buffers travel round the ring by NCCL P2P
   (`batch_isend_irecv` on a side stream, double buffered) while the attention kernel merges the resident shard
   into the fp32 running state (lowbit_attn_fwd_partial); lowbit_attn_finalize normalises at the end.
   Causal attention uses the zig-zag layout (rank r owns chunks r and 2P-1-r) so every rank does the same work,
   and (Q chunk, K chunk) pairs that lie wholly in the future are skipped.

The compute is reached through a small backend object so the host logic (schedule, message packing, stream
protocol) is testable on CPU with gloo: tests inject an oracle-math backend; the product path always uses
`CudaBackend` (hand-written sm_100a kernels through the C ABI) -- there is no CPU fallback in this module.
"""
import math
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from . import _native as N
from . import _tensor as T

LOG2E = 1.44269504


# ----------------------------------------------------------------------------------------------- head sharding
def head_shard(num_qo_heads: int, num_kv_heads: int, world: int, rank: int) -> Tuple[int, int, int, int]:
    """(hq0, hq1, hkv0, hkv1): this rank's q-head and kv-head ranges.  The split is on kv heads so that a
    GQA group stays with its kv head (attn_qk_int8_per_block.py:104,119: `off_h // num_kv_groups`)."""
    assert num_qo_heads % num_kv_heads == 0, "num_qo_heads must be divisible by num_kv_heads"
    if num_kv_heads % world != 0:
        raise ValueError(f"num_kv_heads ({num_kv_heads}) must be divisible by the number of ranks ({world})")
    g = num_qo_heads // num_kv_heads
    per = num_kv_heads // world
    return rank * per * g, (rank + 1) * per * g, rank * per, (rank + 1) * per


def lowbit_fa_head_sharded(q, k, v, fn, world: int, rank: int, tensor_layout: str = "HND", **kwargs):
    """Run `fn` (any lowbit_fa_* entry point) on this rank's head slice of full-size q, k, v.  The slices are
    strided views (NHD: stride_seq = H*D), consumed as such by the kernels; returns the rank's slice of O."""
    hdim = 1 if tensor_layout == "HND" else 2
    qt, kt, vt = T.as_torch(q), T.as_torch(k), T.as_torch(v)
    hq0, hq1, hkv0, hkv1 = head_shard(qt.shape[hdim], kt.shape[hdim], world, rank)
    return fn(qt.narrow(hdim, hq0, hq1 - hq0), kt.narrow(hdim, hkv0, hkv1 - hkv0), vt.narrow(hdim, hkv0, hkv1 - hkv0),
              tensor_layout=tensor_layout, **kwargs)


# ----------------------------------------------------------------------------------------------- ring schedule
@dataclass(frozen=True)
class Chunk:
    offset: int  # global position of the chunk's first token
    length: int


def seq_chunks(n_total: int, world: int, rank: int, zigzag: bool) -> List[Chunk]:
    """Token chunks owned by `rank`: contiguous [r*n/P, (r+1)*n/P), or zig-zag (chunks r and 2P-1-r of 2P) which
    balances causal work across ranks."""
    if zigzag:
        assert n_total % (2 * world) == 0, "zig-zag sharding needs N divisible by 2*world"
        c = n_total // (2 * world)
        return [Chunk(rank * c, c), Chunk((2 * world - 1 - rank) * c, c)]
    assert n_total % world == 0, "sequence sharding needs N divisible by world"
    c = n_total // world
    return [Chunk(rank * c, c)]


def ring_source(rank: int, step: int, world: int) -> int:
    """Rank whose K/V shard is resident on `rank` at ring step `step` (buffers move rank -> rank+1)."""
    return (rank - step) % world


def pair_visible(qc: Chunk, kc: Chunk, causal: bool) -> bool:
    """False when every key of kc lies in the future of every query of qc."""
    return (not causal) or (kc.offset <= qc.offset + qc.length - 1)


# ----------------------------------------------------------------------------------------------- message layout
@dataclass
class Field:
    name: str
    dtype: torch.dtype
    shape: Tuple[int, ...]
    offset: int = 0


class RingMessage:
    """One rank's quantized K/V shard as ONE flat byte buffer (a single NCCL send/recv per ring step), with typed
    views for every field.  Offsets are 256-byte aligned (TMA wants 16; 256 keeps sectors whole)."""

    def __init__(self, fields: List[Field], device):
        off = 0
        self.fields = {}
        for f in fields:
            nbytes = math.prod(f.shape) * torch.empty((), dtype=f.dtype).element_size()
            f.offset = off
            self.fields[f.name] = f
            off = (off + nbytes + 255) // 256 * 256
        self.nbytes = off
        self.flat = torch.empty(self.nbytes, dtype=torch.uint8, device=device)

    def view(self, name: str) -> torch.Tensor:
        f = self.fields[name]
        nbytes = math.prod(f.shape) * torch.empty((), dtype=f.dtype).element_size()
        return self.flat[f.offset:f.offset + nbytes].view(f.dtype).view(f.shape)

    def like(self):
        m = RingMessage.__new__(RingMessage)
        m.fields, m.nbytes = self.fields, self.nbytes
        m.flat = torch.empty_like(self.flat)
        return m


def kv_fields(b, hkv, d, chunks: List[Chunk], tensor_layout, qk, pv) -> List[Field]:
    """Fields of a K/V shard message: per chunk i  k{i} codes, ks{i} scales, v{i}, (vs{i}, vm{i})."""
    fs = []
    for i, c in enumerate(chunks):
        n = c.length
        dk = d // 2 if qk == "int4" else d  # "mixed": the D-byte container (a block uses its first D*bits/8 bytes)
        shp = (lambda dd: (b, hkv, n, dd)) if tensor_layout == "HND" else (lambda dd: (b, n, hkv, dd))
        fs.append(Field(f"k{i}", torch.int8, shp(dk)))
        fs.append(Field(f"ks{i}", torch.float32, (b, hkv, (n + 63) // 64)))
        if qk == "mixed":
            fs.append(Field(f"kb{i}", torch.int32, (b, hkv, (n + 63) // 64)))
        if pv == "fp8":
            npad = (n + 63) // 64 * 64
            fs.append(Field(f"v{i}", torch.float8_e4m3fn, (b, hkv, d, npad) if tensor_layout == "HND" else (b, d, hkv, npad)))
            fs.append(Field(f"vs{i}", torch.float32, (b, hkv, d)))
        else:
            fs.append(Field(f"v{i}", torch.float16, shp(d)))
    return fs


# ----------------------------------------------------------------------------------------------- compute backend
class CudaBackend:
    """The product compute path: sm_100a kernels through the C ABI."""

    def __init__(self, tensor_layout, qk, pv, sm_scale):
        from . import attention as A
        from . import quant as Qz
        self.A, self.Qz = A, Qz
        self.layout, self.qk, self.pv, self.sm_scale = tensor_layout, qk, pv, sm_scale
        self.qk_mode = N.QK_Q8K4 if qk == "int4" else (N.QK_Q8KMIX if qk == "mixed" else N.QK_I8)
        self.k_thresholds = (0.2, 0.05)
        self.pv_mode = N.PV_E4M3 if pv == "fp8" else N.PV_F16

    def k_sum(self, k, seq_dim):
        # exact for fp16 magnitudes met in practice (|k| * N < 2^29 in units of 2^-24): the all-reduced sum is
        # then order independent, and km = fp16(fp32(sum) / N) matches the single-GPU contract
        return k.sum(dim=seq_dim, dtype=torch.float64)

    def quantize_q(self, q_chunk):
        return self.Qz._quant_one(q_chunk, None, 128, 8, False, self.sm_scale * LOG2E, N.QMODE_TRITON, self.layout)

    def quantize_kv(self, k_chunk, v_chunk, km, msg: RingMessage, i: int):
        if self.qk == "mixed":
            msg.view(f"k{i}").zero_()  # unused row tails of the container
            self.Qz.per_block_k_mixed(k_chunk, km, None, self.k_thresholds[0], self.k_thresholds[1], self.layout,
                                      out=(msg.view(f"k{i}"), msg.view(f"ks{i}"), msg.view(f"kb{i}")))
        else:
            bits, pack = (4, True) if self.qk == "int4" else (8, False)
            self.Qz._quant_one(k_chunk, km, 64, bits, pack, 1.0, N.QMODE_TRITON, self.layout,
                               out=(msg.view(f"k{i}"), msg.view(f"ks{i}")))
        if self.pv == "fp8":
            v8, vs, _ = self.Qz.per_channel_fp8(v_chunk, self.layout, smooth_v=False)
            msg.view(f"v{i}").copy_(v8)
            msg.view(f"vs{i}").copy_(vs)
        else:
            msg.view(f"v{i}").copy_(v_chunk)  # bf16 inputs: V travels as fp16 (core.py:307-308)

    def partial(self, state, q_pack, msg: RingMessage, i: int, q_off: int, k_off: int, causal: bool):
        qc, qs = q_pack
        return self.A.forward_partial(state, qc, msg.view(f"k{i}"), msg.view(f"v{i}"), qs, msg.view(f"ks{i}"),
                                      self.layout, causal=causal, q_offset=q_off, k_offset=k_off,
                                      qk_mode=self.qk_mode, pv_mode=self.pv_mode,
                                      v_scale=msg.view(f"vs{i}") if self.pv == "fp8" else None,
                                      kbits=msg.view(f"kb{i}") if self.qk == "mixed" else None)

    def finalize(self, state, q_pack, out_dtype, return_lse):
        return self.A.finalize(state, q_pack[0], self.layout, out_dtype, return_lse)


# ----------------------------------------------------------------------------------------------- the ring
def _exchange(dist, group, send_buf, recv_buf, world, rank):
    nxt, prv = (rank + 1) % world, (rank - 1) % world
    if group is not None:
        nxt, prv = dist.get_global_rank(group, nxt), dist.get_global_rank(group, prv)
    ops = [dist.P2POp(dist.isend, send_buf, nxt, group=group), dist.P2POp(dist.irecv, recv_buf, prv, group=group)]
    return dist.batch_isend_irecv(ops)


def ring_group(timeout_s: float = 60.0, ranks=None, backend: Optional[str] = None):
    """A process group for `ring_attention` whose collectives and P2P transfers time out after `timeout_s` seconds
    (the default NCCL timeout is 10 minutes): with torch's asynchronous NCCL error handling a peer that died mid-ring
    then aborts the communicator instead of hanging the job.  `ranks` = None: all ranks of the world."""
    import datetime

    import torch.distributed as dist
    return dist.new_group(ranks=ranks, timeout=datetime.timedelta(seconds=float(timeout_s)), backend=backend)


def ring_attention(q, k, v, tensor_layout: str = "HND", is_causal: bool = False, sm_scale: Optional[float] = None,
                   smooth_k: bool = True, qk: str = "int4", pv: str = "fp16", zigzag: Optional[bool] = None,
                   return_lse: bool = False, group=None, backend=None, n_total: Optional[int] = None,
                   timings: Optional[dict] = None, check: bool = False):
    """Sequence-parallel low-bit attention over the ranks of `group` (default: the world).

    q, k, v: this rank's shard, [B,H,n_local,D] (HND) or [B,n_local,H,D] (NHD), n_local = N / world, holding the
    rank's chunks back to back in `seq_chunks` order (zig-zag by default for causal).  head_dim 64 or 128.
    Returns o for the local rows (same shape/dtype as q) and, with return_lse, lse [B,Hq,n_local] (natural log,
    of the smoothed scores' softmax -- the (q . km) correction of core.py:344-350 is added like the API does).
    `timings`: a dict that receives this rank's phase times in ms (CUDA events on the compute stream; the call
    synchronises the device to read them): "k_mean", "quantize", per ring step "compute" (attention kernels of the
    resident shard) and "p2p_exposed" (what the step waited for the exchange beyond its own compute), "finalize",
    and "p2p_bytes_per_step" (the flat K/V message one rank sends per step).
    Failure detection: NCCL reports a dead peer or a stuck transfer asynchronously -- after the group's timeout
    (`ring_group(timeout_s)` makes a group with a short one) the watchdog aborts the communicator and the error
    surfaces at the next synchronisation.  `check=True` synchronises the comm stream at the end of every ring step
    and turns such an error into a RuntimeError that names the step and the rank (it serialises host and device once
    per step: a debugging / canary mode, not the default).  Every phase is also an NVTX range (`lowbit.ring.*`)."""
    import torch.distributed as dist
    if tensor_layout not in ("HND", "NHD"):
        raise ValueError(f"Unknown tensor layout: {tensor_layout}")
    if qk not in ("int8", "int4", "mixed") or pv not in ("fp16", "fp8"):
        raise ValueError(f"Unsupported ring formats qk={qk} pv={pv}")
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    zigzag = bool(is_causal) if zigzag is None else zigzag
    seq = 2 if tensor_layout == "HND" else 1
    hdim = 1 if tensor_layout == "HND" else 2
    b, d = q.shape[0], q.shape[-1]
    hq, hkv, n_local = q.shape[hdim], k.shape[hdim], q.shape[seq]
    if d not in (64, 128):
        raise ValueError(f"Unsupported head_dim for the ring path: {d}")
    n_total = n_total or n_local * world
    if sm_scale is None:
        sm_scale = 1.0 / d ** 0.5
    be = backend if backend is not None else CudaBackend(tensor_layout, qk, pv, sm_scale)
    on_cuda = q.device.type == "cuda"
    if backend is None and not on_cuda:
        raise N.LowbitNativeError("ring_attention: tensors must live on a CUDA device (no CPU fallback)")

    import os
    marks = [] if (on_cuda and (timings is not None or os.environ.get("LOWBIT_RING_TIMING"))) else None  # (label, event)

    nvtx_open = [False]

    def mark(label):
        if on_cuda:  # NVTX: one range per phase, closed by the next mark (no-ops without a profiler attached)
            if nvtx_open[0]:
                torch.cuda.nvtx.range_pop()
            nvtx_open[0] = label != "finalize"
            if nvtx_open[0]:
                torch.cuda.nvtx.range_push("lowbit.ring after " + label)
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(q.device))
            marks.append((label, ev))

    mark("start")
    my_chunks = seq_chunks(n_total, world, rank, zigzag)
    assert sum(c.length for c in my_chunks) == n_local, "local shard length does not match N / world"

    def local_views(x):
        out, pos = [], 0
        for c in my_chunks:
            out.append(x.narrow(seq, pos, c.length))
            pos += c.length
        return out

    # global K mean: one all-reduce of [B,Hkv,D] fp64 sums
    km = None
    if smooth_k:
        ksum = be.k_sum(k, seq)
        dist.all_reduce(ksum, group=group)
        km = (ksum.to(torch.float32) / torch.tensor(float(n_total), dtype=torch.float32, device=k.device)).to(k.dtype)
        km = km.contiguous()  # [B,Hkv,D] (seq dim reduced away in either layout)
    mark("k mean (sum + all-reduce)")

    # quantize once: Q chunks stay resident, the K/V shard goes into the flat ring message
    q_packs = [be.quantize_q(x) for x in local_views(q)]
    msg = RingMessage(kv_fields(b, hkv, d, my_chunks, tensor_layout, qk, pv), q.device)
    for i, (kx, vx) in enumerate(zip(local_views(k), local_views(v))):
        be.quantize_kv(kx, vx.to(torch.float16) if (pv == "fp16" and vx.dtype != torch.float16) else vx, km, msg, i)
    bufs = [msg, msg.like()]
    mark("quantize q, k, v")

    compute_stream = torch.cuda.current_stream(q.device) if on_cuda else None
    comm_stream = torch.cuda.Stream(q.device) if on_cuda else None
    states = [None] * len(my_chunks)
    cur = 0
    for step in range(world):
        src = ring_source(rank, step, world)
        works = None
        if step + 1 < world:
            if on_cuda:
                comm_stream.wait_stream(compute_stream)  # bufs[cur] is complete; bufs[1-cur]'s readers are done
                with torch.cuda.stream(comm_stream):
                    works = _exchange(dist, group, bufs[cur].flat, bufs[1 - cur].flat, world, rank)
            else:
                works = _exchange(dist, group, bufs[cur].flat, bufs[1 - cur].flat, world, rank)
        src_chunks = seq_chunks(n_total, world, src, zigzag)
        for qi, qc in enumerate(my_chunks):
            for ki, kc in enumerate(src_chunks):
                if pair_visible(qc, kc, is_causal):
                    states[qi] = be.partial(states[qi], q_packs[qi], bufs[cur], ki, qc.offset, kc.offset, bool(is_causal))
        mark(f"compute {step}")
        if works is not None:
            if on_cuda:
                with torch.cuda.stream(comm_stream):
                    for w in works:
                        w.wait()
                if check:
                    try:
                        comm_stream.synchronize()
                    except RuntimeError as e:  # NCCL watchdog abort / CUDA error of the exchange
                        raise RuntimeError(f"ring_attention: exchange of step {step} failed on rank {rank} "
                                           f"(peers {(rank - 1) % world} -> {rank} -> {(rank + 1) % world}): {e}") from e
                compute_stream.wait_stream(comm_stream)
            else:
                for w in works:
                    w.wait()
        cur ^= 1
        mark(f"step {step}")

    outs, lses = [], []
    for qi in range(len(my_chunks)):
        o_i, lse_i = be.finalize(states[qi], q_packs[qi], q.dtype, return_lse)
        outs.append(o_i)
        lses.append(lse_i)
    o = outs[0] if len(outs) == 1 else torch.cat(outs, dim=seq)
    mark("finalize")
    if marks is not None:
        torch.cuda.synchronize(q.device)
        spans = [(b_[0], a_[1].elapsed_time(b_[1])) for a_, b_ in zip(marks, marks[1:])]
        if timings is not None:
            timings.clear()
            timings.update({"k_mean": 0.0, "quantize": 0.0, "compute": [], "p2p_exposed": [], "finalize": 0.0,
                            "p2p_bytes_per_step": int(msg.nbytes), "steps": world})
            for label, ms in spans:
                if label.startswith("k mean"):
                    timings["k_mean"] = ms
                elif label.startswith("quantize"):
                    timings["quantize"] = ms
                elif label.startswith("compute"):
                    timings["compute"].append(ms)
                elif label.startswith("step"):
                    timings["p2p_exposed"].append(ms)
                elif label == "finalize":
                    timings["finalize"] = ms
        if rank == 0 and os.environ.get("LOWBIT_RING_TIMING"):
            print("ring timing (ms): " + ", ".join(f"{lb} {ms:.3f}" for lb, ms in spans), flush=True)
    if not return_lse:
        return o
    lse2 = lses[0] if len(lses) == 1 else torch.cat(lses, dim=2)
    lse = lse2 / LOG2E
    if smooth_k:
        qh = q if tensor_layout == "HND" else q.permute(0, 2, 1, 3)
        kmh = km.repeat_interleave(hq // hkv, dim=1) if hq != hkv else km
        corr = torch.einsum("bhnd,bhd->bhn", qh.float(), kmh.float()).to(q.dtype).float()
        lse = lse + corr * sm_scale
    return o, lse
