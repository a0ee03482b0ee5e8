#!/usr/bin/env python
"""bench.py -- headline benchmark of the low-bit attention hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch: K mean -> per-block INT8 quantize of Q and K (K smoothing
fused) -> fused INT8-QK / FP16-PV attention, i.e. one call of lowbit_fa_qk_int8_pv_fp16_triton.
Workload at N=1: BASELINE config 2 (B4 H32 N4096 D64 HND non-causal, randn fp16, seed 0).
metric = attention TOPS = 4*B*H*Nq*Nk*D / latency (utils/benchmark.py:212-214 of the reference).

  value          whole hot path (quantize + attention), inputs resident in HBM, CUDA events, max over ranks
  attn_only      the attention kernel alone (how the reference's published numbers are measured)
  e2e            the same operator from HOST pinned buffers through lowbit_fa_host: H2D of q,k,v + hot path + D2H of o
                 inside the timed region, pipelined over (batch, head-group) chunks; serial_ms = the unpipelined time
  roofline       dominant kernel (attention): algorithmic FLOP per launch / its mean CUDA-event duration inside the
                 timed steps, against the measured dense bf16 peak of MEASURED_PEAKS.json
  cpu_baseline   oracle port of the reference's pure-Paddle quantize-and-attend math on the host cores, bounded sample
N > 1: every rank runs the same workload on its own GPU (batch x head units are independent: no collective);
value = units of all ranks / max-over-ranks time; "scaling": "weak".

--impl reference: the reference's CPU path (oracle port; Paddle/Triton cannot run the reference on this box's
CPU other than through the interpreter) on a bounded sample with all host threads.
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B, Hq, Hkv, N, D, layout, causal, description)        -- INT8 QK + FP16 PV, every rank the same work (weak)
    "c2": (4, 32, 32, 4096, 64, "HND", False, "BASELINE config 2: INT8 QK + FP16 PV, HND, B4 H32 N4096 D64 non-causal"),
    "c2c": (4, 32, 32, 4096, 64, "HND", True, "config 2 shape, causal"),
    "c3_8k": (4, 32, 32, 8192, 128, "HND", True, "config 3 shape (D128 causal 8K), INT8 QK + FP16 PV"),
    "c4": (2, 48, 48, 17776, 64, "NHD", False, "config 4 shape: CogVideoX-5B B2 H48 N17776 D64 NHD"),
}
# the other BASELINE configs: (B, Hq, Hkv, N, D, layout, causal, qk, pv, partition, description)
#   partition "replicate": every rank the whole workload (weak); "heads": kv-head slices, no collective (strong);
#   "ring": sequence-parallel NCCL P2P ring of quantized K/V (strong)
EXTRA = {
    "c3q_8k": (4, 32, 32, 8192, 128, "HND", True, "int4", "fp8", "replicate", "BASELINE config 3: INT4 QK + FP8 PV, HND, B4 H32 D128 causal N=8K"),
    "c3q_16k": (2, 32, 32, 16384, 128, "HND", True, "int4", "fp8", "replicate", "BASELINE config 3: INT4 QK + FP8 PV, HND, B2 H32 D128 causal N=16K"),
    "c3q_32k": (1, 32, 32, 32768, 128, "HND", True, "int4", "fp8", "replicate", "BASELINE config 3: INT4 QK + FP8 PV, HND, B1 H32 D128 causal N=32K"),
    "c4s": (2, 48, 48, 17776, 64, "NHD", False, "q8k4", "fp16", "heads", "BASELINE config 4: q_int8/k_int4, NHD, CogVideoX-5B B2 H48 N17776 D64, head-sharded"),
    "c5": (1, 32, 32, 131072, 128, "HND", True, "int4", "fp16", "ring", "BASELINE config 5 (INT4 K): B1 H32 N128K D128 causal, sequence-parallel ring of quantized K/V"),
    "c5dyn": (1, 32, 32, 131072, 128, "HND", True, "mixed", "fp16", "ring", "BASELINE config 5: dynamic INT4/INT2 K bit allocation (per 64-key block), B1 H32 N128K D128 causal, sequence-parallel ring of quantized K/V"),
    "c5f8": (1, 32, 32, 131072, 128, "HND", True, "int4", "fp8", "ring", "BASELINE config 5 (INT4 K, FP8 V): B1 H32 N128K D128 causal, sequence-parallel ring"),
}
BASELINE_MD_TOPS = 199.5  # INT8 non-causal B4 H32 D64 N=4096, attention kernel only, hardware unstated


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the attention launch at C2 from the committed `ncu --set full`
    summary (profiles/r1_attn_c2_ncu_summary.json, latest kernel version listed there)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_attn_c2_ncu_summary.json")) as f:
            last = list(json.load(f).values())[-1]
        mb = lambda s: float(s.split()[0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3}[s.split()[1]]
        return mb(last["dram__bytes_read.sum"]) + mb(last["dram__bytes_write.sum"])
    except Exception:
        return None


def mufu_bound(B, Hq, N, causal, attn_ms, sm_mhz, sms=148):
    """exp2 evaluations per attention launch against the XU pipe's 16 MUFU.EX2 per clock per SM at the SM clock sampled
    during the run (1965 MHz when the sample is missing)."""
    exps = float(B) * Hq * N * N / (2 if causal else 1)
    rate = sms * 16 * (sm_mhz or 1965) * 1e6
    floor_ms = exps / rate * 1e3
    return {"exp2_per_launch": exps, "peak_exp2_per_s": rate, "floor_ms": floor_ms, "frac": floor_ms / attn_ms}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops"]), float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json bf16 burst)"
    except Exception:
        return 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def bind_near_gpu(index):
    """Pin this process to the CPU cores NVML reports as local to GPU `index` (same NUMA node / PCIe root), BEFORE the
    pinned host buffers of the e2e leg are allocated, so that first-touch places them next to the GPU.  Without it the
    H2D rate of the e2e leg varies 2x from run to run on the multi-socket GPU boxes."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def dist_setup(ngpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def cpu_baseline_run(wl, sample_heads=32, reps=1, causal=None):
    """Time the oracle port (reference pure-Paddle math restated on torch-CPU) on (1 batch x sample_heads heads)
    of the workload with all host threads.  Returns (tops, cores, sample description, seconds)."""
    from oracle import attention as OA
    B, Hq, Hkv, N, D, layout, caus, _ = WORKLOADS[wl]
    causal = caus if causal is None else causal
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    g = torch.Generator().manual_seed(0)
    shp = (1, sample_heads, N, D) if layout == "HND" else (1, N, sample_heads, D)
    q, k, v = (torch.randn(shp, generator=g).half() for _ in range(3))
    OA.cpu_quantize_and_attend(q[:, :1] if layout == "HND" else q[:, :, :1], k[:, :1] if layout == "HND" else k[:, :, :1],
                               v[:, :1] if layout == "HND" else v[:, :, :1], layout, causal)  # warm
    t0 = time.perf_counter()
    for _ in range(reps):
        for h0 in range(0, sample_heads, 4):  # 4 heads at a time bounds the fp32 score matrix to ~1 GiB
            sl = (slice(None), slice(h0, h0 + 4)) if layout == "HND" else (slice(None), slice(None), slice(h0, h0 + 4))
            OA.cpu_quantize_and_attend(q[sl], k[sl], v[sl], layout, causal)
    dt = (time.perf_counter() - t0) / reps
    ops = 4.0 * sample_heads * N * N * D / (2 if causal else 1)
    return ops / dt / 1e12, cores, f"1 batch x {sample_heads} heads of {wl} (N={N}, D={D}), fp32 math, {reps} rep(s)", dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    B, Hq, Hkv, N, D, layout, causal, desc = WORKLOADS[wl]
    heads = 32 if N <= 4096 else (8 if N <= 8192 else 4)
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_baseline_run(wl, sample_heads=4)
    vals = []
    t0 = time.perf_counter()
    for _ in range(max(1, min(args.steps, 5))):
        tops, cores, sample, dt = cpu_baseline_run(wl, sample_heads=heads)
        vals.append((tops, dt))
        if time.perf_counter() - t0 > 120:
            break
    tops = sum(v[0] for v in vals) / len(vals)
    ms = sum(v[1] for v in vals) / len(vals) * 1e3
    line = {
        "impl": "reference", "metric": "attention TOPS (4*B*H*N^2*D / latency), quantize + attention", "value": tops,
        "unit": "TOPS", "n_gpus": args.gpus, "steps": len(vals), "warmup": 1, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8 QK / fp16 PV (fp32 on CPU)",
        "data": "synthetic randn fp16 seed 0",
        "config": {"workload": desc, "sample": sample},
        "cpu_baseline": {"value": tops, "unit": "TOPS", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": tops, "unit": "TOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



def run_extra(args):
    """BASELINE configs 3-5 (not the driver's default line): same metric, API-level step, CUDA events, max over ranks."""
    rank, world, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    import lowbit_quant_fa2_paddle_b200 as L
    from lowbit_quant_fa2_paddle_b200 import parallel as P
    B, Hq, Hkv, N, D, layout, causal, qk, pv, part, desc = EXTRA[args.workload]
    W, K = max(args.warmup, 3), args.steps
    torch.manual_seed(0)
    seq = 2 if layout == "HND" else 1
    shp = lambda h, n=N: (B, h, n, D) if layout == "HND" else (B, n, h, D)
    ops_total = 4.0 * B * Hq * N * N * D / (2 if causal else 1)
    if qk == "mixed":
        fn = L.lowbit_fa_q_int8_k_dynamic
    elif pv == "fp8":
        fn = L.lowbit_fa_qk_int4_pv_fp8 if qk != "int8" else L.lowbit_fa_qk_int8_pv_fp8_cuda
    else:
        fn = {"int8": L.lowbit_fa_qk_int8_pv_fp16_triton, "int4": L.lowbit_fa_qk_int4_pv_fp16_triton,
              "q8k4": L.lowbit_fa_q_int8_k_int4_pv_fp16}[qk]

    def dyn(k):
        """dynamic-K workloads: every other 64-key block at 0.2x magnitude, so that the block statistic
        max|k|/127 puts it in the INT2 class (<= 0.0125) and the rest in the INT4 class (randn: ~0.035)"""
        if qk != "mixed":
            return k
        n = k.shape[seq]
        w = torch.where((torch.arange(n, device=dev) // 64) % 2 == 1, 0.2, 1.0).to(k.dtype)
        return k * (w.view(1, 1, n, 1) if layout == "HND" else w.view(1, n, 1, 1))
    if part == "ring" and world > 1:
        n_loc = N // world
        q, k, v = (torch.randn(shp(h, n_loc), dtype=torch.float16, device=dev) for h in (Hq, Hkv, Hkv))
        k = dyn(k)
        step = lambda: P.ring_attention(q, k, v, tensor_layout=layout, is_causal=causal,
                                        qk=qk if qk in ("int8", "mixed") else "int4", pv=pv)
        scaling, units = "strong", ops_total
    elif part == "heads" and world > 1:
        q, k, v = (torch.randn(shp(h), dtype=torch.float16, device=dev) for h in (Hq, Hkv, Hkv))
        step = lambda: P.lowbit_fa_head_sharded(q, k, v, fn, world, rank, tensor_layout=layout, is_causal=causal)
        scaling, units = "strong", ops_total
    else:
        q, k, v = (torch.randn(shp(h), dtype=torch.float16, device=dev) for h in (Hq, Hkv, Hkv))
        k = dyn(k)
        step = lambda: fn(q, k, v, tensor_layout=layout, is_causal=causal)
        scaling, units = ("strong", ops_total) if part != "replicate" else ("weak", ops_total * world)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(W):
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        step()
    e1.record(stream)
    barrier()
    sampler.stop_flag = True
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / K
    if rank == 0:
        peak_tf, _, peak_src = peaks()
        value = units / (ms * 1e-3) / 1e12
        per_gpu = value / world
        line = {"metric": "attention TOPS (4*B*H*N^2*D / latency), quantize + attention", "value": value, "unit": "TOPS",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": f"{qk} QK (int32 acc) / {pv} PV (fp32 acc)",
                "data": "synthetic randn fp16 seed 0" + ("; every other 64-key block of K at 0.2x (INT2 class)" if qk == "mixed" else ""),
                "config": {"workload": desc, "partition": part, "l2": "inputs larger than the 126 MB L2, no flush"},
                "roofline": {"bound": "tensor", "kernel": "attn_fwd_kernel (whole step timed: quantize + attention"
                             + (" + ring exchange" if part == "ring" else "") + ")", "achieved": per_gpu, "peak": peak_tf,
                             "unit": "TFLOP/s", "frac": per_gpu / peak_tf, "traffic": None, "peak_source": peak_src},
                "clocks": sampler.summary(), "gpu_launches": None}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + sorted(EXTRA))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.workload in EXTRA:
        if args.impl == "reference":
            raise SystemExit("--impl reference is defined for the INT8/FP16 workloads")
        return run_extra(args)
    if args.impl == "reference":
        return run_reference(args)

    rank, world, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    affinity = bind_near_gpu(local)
    import lowbit_quant_fa2_paddle_b200 as L
    from lowbit_quant_fa2_paddle_b200 import _native
    _native.lib()

    B, Hq, Hkv, N, D, layout, causal, desc = WORKLOADS[args.workload]
    W = max(args.warmup, 3)
    K = args.steps
    torch.manual_seed(rank)  # seed 0 on rank 0
    shp = lambda h: (B, h, N, D) if layout == "HND" else (B, N, h, D)
    q = torch.randn(shp(Hq), dtype=torch.float16, device=dev)
    k = torch.randn(shp(Hkv), dtype=torch.float16, device=dev)
    v = torch.randn(shp(Hkv), dtype=torch.float16, device=dev)
    ops = 4.0 * B * Hq * N * N * D / (2 if causal else 1)
    stream = torch.cuda.current_stream(dev)

    # ---- the step, with CUDA events around the dominant kernel (attention) inside it ----
    from lowbit_quant_fa2_paddle_b200 import attention as A
    from lowbit_quant_fa2_paddle_b200 import quant as Qz
    attn_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sm_scale = 1.0 / D ** 0.5

    def step(i=None):
        """== lowbit_fa_qk_int8_pv_fp16_triton(q,k,v,...) with event marks around the attention launch."""
        qc, qs, kc, ks, _ = Qz.smooth_and_quantize(q, k, True, sm_scale, layout, 8, 8, False, "triton")
        if i is not None:
            attn_ev[i][0].record(stream)
        o, _ = A._forward(qc, kc, v, qs, ks, layout, torch.float16, False, causal)
        if i is not None:
            attn_ev[i][1].record(stream)
        return o

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(W):
        step()
    # parity spot-check of the public API against the step decomposition above (bit-identical)
    o_api = L.lowbit_fa_qk_int8_pv_fp16_triton(q, k, v, tensor_layout=layout, is_causal=causal)
    assert torch.equal(o_api, step()), "bench step differs from the public API call"
    del o_api

    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(K):
        step(i)
    e1.record(stream)
    barrier()
    sampler.stop_flag = True
    total_ms = e0.elapsed_time(e1)
    attn_ms = sum(a.elapsed_time(b) for a, b in attn_ev) / K

    # ---- attention kernel alone (the reference's published style: quantization outside the timed region) ----
    km = Qz.k_mean(k, layout)
    qc, qs, kc, ks = Qz.per_block_int8(q, k, km=km, sm_scale=sm_scale, tensor_layout=layout)
    for _ in range(3):
        A._forward(qc, kc, v, qs, ks, layout, torch.float16, False, causal)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    KA = max(10, min(K, 50))
    a0.record(stream)
    for _ in range(KA):
        A._forward(qc, kc, v, qs, ks, layout, torch.float16, False, causal)
    a1.record(stream)
    torch.cuda.synchronize(dev)
    attn_alone_ms = a0.elapsed_time(a1) / KA

    # ---- end to end from host pinned memory (H2D q,k,v + hot path + D2H o) ----
    hq_, hk_, hv_ = (t.cpu().pin_memory() for t in (q, k, v))
    ho = torch.empty(q.shape, dtype=torch.float16).pin_memory()
    KE = max(3, min(K, 20))

    def e2e_step():
        """The user-facing host entry point: pinned host q,k,v -> (H2D | quantize + attention | D2H, pipelined over
        (batch, head-group) chunks on three streams) -> pinned host o."""
        L.lowbit_fa_host(hq_, hk_, hv_, out=ho, tensor_layout=layout, is_causal=causal)

    def e2e_serial_step():
        """Same bytes, no overlap (copy in, one operator call, copy out): reported as e2e.serial_ms for context."""
        dq = hq_.to(dev, non_blocking=True)
        dk = hk_.to(dev, non_blocking=True)
        dv = hv_.to(dev, non_blocking=True)
        o = L.lowbit_fa_qk_int8_pv_fp16_triton(dq, dk, dv, tensor_layout=layout, is_causal=causal)
        ho.copy_(o, non_blocking=True)

    e2e_serial_step()
    torch.cuda.synchronize(dev)
    ho_serial = ho.clone()
    y0, y1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    y0.record(stream)
    for _ in range(3):
        e2e_serial_step()
    y1.record(stream)
    torch.cuda.synchronize(dev)
    e2e_serial_ms = y0.elapsed_time(y1) / 3
    ho.zero_()
    for _ in range(3):
        e2e_step()
    barrier()
    xe = [torch.cuda.Event(enable_timing=True) for _ in range(KE + 1)]
    xe[0].record(stream)
    for i in range(KE):
        e2e_step()
        xe[i + 1].record(stream)  # lowbit_fa_host leaves the caller's stream waiting for the last copy-out
    barrier()
    per = sorted(xe[i].elapsed_time(xe[i + 1]) for i in range(KE))
    e2e_ms = xe[0].elapsed_time(xe[KE]) / KE
    e2e_median_ms = per[KE // 2]
    print("e2e per-step ms (sorted):", " ".join(f"{t:.2f}" for t in per), file=sys.stderr)
    assert torch.equal(ho, ho_serial), "pipelined host entry point differs from the serial call"

    # ---- max over ranks ----
    e2e_mean_ms, e2e_max_ms = e2e_ms, per[-1]
    e2e_ms = e2e_median_ms  # robust: 1 step in ~30 is hit by a 50-100 ms host/driver stall
    t = torch.tensor([total_ms, attn_ms, attn_alone_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, attn_ms, attn_alone_ms, e2e_ms = t.tolist()
    ms_per_step = total_ms / K

    if rank == 0:
        peak_tf, peak_bw, peak_src = peaks()
        value = world * ops / (ms_per_step * 1e-3) / 1e12
        achieved = ops / (attn_ms * 1e-3) / 1e12
        line = {
            "metric": "attention TOPS (4*B*H*N^2*D / latency), quantize + attention", "value": value, "unit": "TOPS",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": value / BASELINE_MD_TOPS if args.workload == "c2" else None,
            "dtype": "int8 QK (int32 acc) / fp16 PV (fp32 acc)", "data": "synthetic randn fp16 seed 0",
            "config": {"workload": desc, "per_gpu": True, "l2": "working set ~320 MiB (q,k,v,codes,o) > 126 MB L2, no flush",
                       "smooth_k": True, "quantization_backend": "triton (Q1 rounding)",
                       "vs_baseline_note": "BASELINE.md 199.5 TFLOP/s is attention-kernel-only on unstated hardware; value includes quantization"},
            "attn_only": {"value": world * ops / (attn_alone_ms * 1e-3) / 1e12, "unit": "TOPS", "ms": attn_alone_ms},
            "e2e": {"value": world * ops / (e2e_ms * 1e-3) / 1e12, "unit": "TOPS", "ms": e2e_ms,
                    "api": "lowbit_fa_host (pinned host q,k,v -> pinned host o; 9 chunks on 3 streams, replayed as one CUDA graph)",
                    "serial_ms": e2e_serial_ms, "statistic": "median of per-step CUDA-event times", "mean_ms": e2e_mean_ms,
                    "max_ms": e2e_max_ms, "steps": KE, "host_cpus_bound": affinity,
                    "h2d_bytes_per_step": int(hq_.numel() * 2 + hk_.numel() * 2 + hv_.numel() * 2),
                    "d2h_bytes_per_step": int(ho.numel() * 2)},
            "gpu_launches": 5 * K,
            "roofline": {"bound": "tensor", "kernel": "attn_fwd_kernel", "achieved": achieved, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "traffic": ncu_traffic_bytes() if args.workload == "c2" else None, "peak_source": peak_src,
                         "algorithmic_flop_per_launch": ops, "ms_per_launch": attn_ms,
                         # the co-bound that actually binds at D=64: one MUFU.EX2 per score,
                         # 16 per clock per SM
                         "mufu_co_bound": mufu_bound(B, Hq, N, causal, attn_ms, (sampler.summary() or {}).get("sm_mhz"))},
            "clocks": sampler.summary(),
        }
        if not args.no_cpu_baseline:
            tops, cores, sample, _ = cpu_baseline_run(args.workload, sample_heads=32 if N <= 4096 else (8 if N <= 8192 else 4))
            line["cpu_baseline"] = {"value": tops, "unit": "TOPS", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
