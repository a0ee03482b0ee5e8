#!/usr/bin/env python
"""bench.py -- headline benchmark of the low-bit attention hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

metric = attention TOPS = 4*B*H*Nq*Nk*D / latency (/2 causal), utils/benchmark.py:212-214 of the reference.
One "step" = one pass of the hot path over one batch: K mean -> per-block quantize of Q and K (K smoothing fused) ->
fused low-bit attention, i.e. one call of the operator the workload names.

Default workloads (what the driver runs):
  N = 1   BASELINE config 2: INT8 QK + FP16 PV, HND, B4 H32 N4096 D64 non-causal (lowbit_fa_qk_int8_pv_fp16_triton);
          the other single-GPU configs ride along as `other_configs` (one short timing each).
  N > 1   BASELINE config 4, the north_star's head-sharded partition: q_int8 / k_int4, NHD, CogVideoX-5B shape B2 H48
          N17776 D64, kv heads split over the ranks, NO collective on the data path, "scaling": "strong"
          (value = ops of the whole workload / max-over-ranks step time); `strong_scaling.single_gpu` is the same
          workload on ONE GPU measured in the same run (rank 0), so that the speed-up is self-contained.
          `ring` carries BASELINE config 5 (B1 H32 N128K D128 causal, sequence-parallel NCCL P2P ring of quantized
          K/V; INT4 K and dynamic INT4/INT2 K) with per-phase times and the bytes one rank sends per step.

Keys of the JSON line:
  value          whole hot path (quantize + attention) through the public operator call, inputs resident in HBM, CUDA
                 events around the K steps, max over ranks
  attn_only      the attention kernel alone (how the reference's published numbers are measured)
  e2e            the same operator from HOST pinned buffers through lowbit_fa_host: H2D of q,k,v + hot path + D2H of o
                 inside the timed region, pipelined over (batch, head-group) chunks; `h2d_gbs_per_rank` = every rank's
                 plain pinned-host -> device copy rate with all ranks copying at once (what bounds e2e at N > 1)
  roofline       dominant kernel (attention): algorithmic op per launch / its mean CUDA-event duration over a timed
                 region of the same K steps run a second time with an event pair around the attention launch
                 (`instrumented_ms_per_step`; the headline steps are the plain public call and carry no events).  `frac` is against the measured dense bf16 peak (MEASURED_PEAKS.json, the contract's number);
                 `t_min` is SURVEY 8d's bound with every term measured: max(QK ops / INT8 peak + PV ops / FP16|FP8 peak,
                 exp2 count / MUFU rate at the sampled SM clock, algorithmic bytes / HBM peak), `bound` names the binding
                 term and `frac_of_t_min` = t_min / measured time.  `traffic` comes from the committed `ncu --set full`
                 capture of the same kernel and shape (profiles/), not from this run.
  cpu_baseline   oracle port of the reference's pure-Paddle quantize-and-attend math on the host cores (N = 1 only)
  reference_gpu  the reference's own Triton kernels JIT-compiled for this GPU (staged under oracle/_ref), N = 1 only

--impl reference: the reference's CPU path (oracle port of its pure-Paddle math; Paddle is not installed and its Triton
kernels have no CPU path other than the interpreter) with all host threads, on the same config as our arm at that N,
each step a bounded sample of (batch, head) units scaled to the workload (`cpu_baseline.sample` says which).
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

# name: B, Hq, Hkv, N, D, layout, causal, qk, pv, partition, description
#   partition "replicate": every rank the whole workload (weak); "heads": kv-head slices, no collective (strong);
#   "ring": sequence-parallel NCCL P2P ring of quantized K/V (strong)
WL = {
    "c2": (4, 32, 32, 4096, 64, "HND", False, "int8", "fp16", "replicate", "BASELINE config 2: INT8 QK + FP16 PV, HND, B4 H32 N4096 D64 non-causal"),
    "c2c": (4, 32, 32, 4096, 64, "HND", True, "int8", "fp16", "replicate", "config 2 shape, causal"),
    "c3_8k": (4, 32, 32, 8192, 128, "HND", True, "int8", "fp16", "replicate", "config 3 shape (D128 causal 8K), INT8 QK + FP16 PV"),
    "c3q_8k": (4, 32, 32, 8192, 128, "HND", True, "int4", "fp8", "replicate", "BASELINE config 3: INT4 QK + FP8 PV, HND, B4 H32 D128 causal N=8K"),
    "c3q_16k": (2, 32, 32, 16384, 128, "HND", True, "int4", "fp8", "replicate", "BASELINE config 3: INT4 QK + FP8 PV, HND, B2 H32 D128 causal N=16K"),
    "c3q_32k": (1, 32, 32, 32768, 128, "HND", True, "int4", "fp8", "replicate", "BASELINE config 3: INT4 QK + FP8 PV, HND, B1 H32 D128 causal N=32K"),
    "c4": (2, 48, 48, 17776, 64, "NHD", False, "int8", "fp16", "replicate", "config 4 shape, INT8 QK: CogVideoX-5B B2 H48 N17776 D64 NHD"),
    "c4s": (2, 48, 48, 17776, 64, "NHD", False, "q8k4", "fp16", "heads", "BASELINE config 4: q_int8/k_int4, NHD, CogVideoX-5B B2 H48 N17776 D64, head-sharded"),
    "c5": (1, 32, 32, 131072, 128, "HND", True, "int4", "fp16", "ring", "BASELINE config 5 (INT4 K): B1 H32 N128K D128 causal, sequence-parallel ring of quantized K/V"),
    "c5dyn": (1, 32, 32, 131072, 128, "HND", True, "mixed", "fp16", "ring", "BASELINE config 5: dynamic INT4/INT2 K bit allocation (per 64-key block), B1 H32 N128K D128 causal, sequence-parallel ring of quantized K/V"),
    "c5f8": (1, 32, 32, 131072, 128, "HND", True, "int4", "fp8", "ring", "BASELINE config 5 (INT4 K, FP8 V): B1 H32 N128K D128 causal, sequence-parallel ring"),
}
BASELINE_MD_TOPS = 199.5  # INT8 non-causal B4 H32 D64 N=4096, attention kernel only, hardware unstated
METRIC = "attention TOPS (4*B*H*N^2*D / latency), quantize + attention"
NCU_SUMMARY = os.path.join(ROOT, "profiles", "r2_attn_c2_ncu_summary.json")


# ------------------------------------------------------------------------------------------------ peaks and bounds
def peaks():
    """Measured denominators: MEASURED_PEAKS.json (driver: HBM copy, bf16 GEMM) and profiles/peaks_int8_fp8.json
    (tools/measure_peaks.py on this pool's B200: INT8 / FP8 / FP16 library GEMMs, the same recipe)."""
    p = {"bf16": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)", "int8": 4500.0, "fp8": 4500.0,
         "fp16": 2250.0, "lowbit_source": "nominal dense (B200_PROFILING.md table)"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update(bf16=float(m["bf16_tflops"]), hbm_gbs=float(m["hbm_gbs"]), source="measured (MEASURED_PEAKS.json bf16 burst)")
    except Exception:
        pass
    try:
        with open(os.path.join(ROOT, "profiles", "peaks_int8_fp8.json")) as f:
            m = json.load(f)
        p.update(int8=float(m["int8"]["burst_tops"]), fp8=float(m["fp8_e4m3"]["burst_tops"]), fp16=float(m["fp16"]["burst_tops"]),
                 lowbit_source="measured (profiles/peaks_int8_fp8.json: cuBLASLt s8, e4m3 and fp16 GEMM 8192^3, burst)")
    except Exception:
        pass
    return p


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the attention launch at config 2 from the committed
    `ncu --set full` summary (latest kernel version listed there)."""
    try:
        with open(NCU_SUMMARY) as f:
            last = list(json.load(f).values())[-1]
        mb = lambda s: float(s.split()[0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3}[s.split()[1]]
        return mb(last["dram__bytes_read.sum"]) + mb(last["dram__bytes_write.sum"])
    except Exception:
        return None


def roofline(kernel, B, Hq, Hkv, N, D, causal, qk, pv, attn_ms, sm_mhz, traffic=None, sms=148):
    """The contract's roofline object plus SURVEY 8d's t_min with measured denominators."""
    pk = peaks()
    div = 2.0 if causal else 1.0
    ops = 4.0 * B * Hq * N * N * D / div
    exps = float(B) * Hq * N * N / div
    kbytes = {"int8": 1.0, "int4": 0.5, "q8k4": 0.5, "mixed": 0.75}[qk]
    vbytes = 2.0 if pv == "fp16" else 1.0
    # algorithmic bytes of the attention launch: Q codes + K codes + V in, O out (scales are noise)
    abytes = B * N * D * (Hq * 1.0 + Hkv * kbytes + Hkv * vbytes + Hq * 2.0)
    t_tensor = (ops / 2) / (pk["int8"] * 1e12) + (ops / 2) / ((pk["fp16"] if pv == "fp16" else pk["fp8"]) * 1e12)
    mufu_rate = sms * 16 * (sm_mhz or 1965) * 1e6
    t_mufu = exps / mufu_rate
    t_hbm = abytes / (pk["hbm_gbs"] * 1e9)
    terms = {"tensor": t_tensor, "mufu": t_mufu, "hbm": t_hbm}
    bound = max(terms, key=terms.get)
    t_min = terms[bound]
    achieved = ops / (attn_ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": kernel, "achieved": achieved, "peak": pk["bf16"], "unit": "TFLOP/s",
            "frac": achieved / pk["bf16"], "traffic": traffic,
            "traffic_source": (os.path.relpath(NCU_SUMMARY, ROOT) + " (ncu --set full capture of the same kernel and shape, not measured in this run)") if traffic else None,
            "peak_source": pk["source"], "algorithmic_op_per_launch": ops, "algorithmic_bytes_per_launch": abytes,
            "ms_per_launch": attn_ms,
            "t_min": {"ms": t_min * 1e3, "binding": bound, "frac_of_t_min": t_min * 1e3 / attn_ms,
                      "tensor_ms": t_tensor * 1e3, "mufu_ms": t_mufu * 1e3, "hbm_ms": t_hbm * 1e3,
                      "int8_peak_tops": pk["int8"], "pv_peak_tops": pk["fp16"] if pv == "fp16" else pk["fp8"],
                      "mufu_exp2_per_s": mufu_rate, "sm_mhz_used": sm_mhz or 1965, "hbm_gbs": pk["hbm_gbs"],
                      "peaks_source": pk["lowbit_source"],
                      "note": "a kernel that takes exp2 off the MUFU pipe (2 of 8 score pairs here) can beat mufu_ms"}}


def cool_down(dev, seconds=1.5):
    """Idle between timed regions: back-to-back regions of ~100 steps were seen to run the second one at a lower
    clock (attention 0.61 -> 0.66 ms); every region gets its own clock samples as well."""
    torch.cuda.synchronize(dev)
    time.sleep(seconds)


def instrumented(instr_ms, attn_ms, ms_per_step, clocks=None):
    """Where `roofline.ms_per_launch` was measured: a second timed region of the same K steps with a CUDA-event pair
    around the attention launch; the headline steps carry no events."""
    return {"measured_in": "instrumented pass: the same K steps, operator cut at the attention launch, CUDA-event pair around it "
                           "on the launching stream; `ms_per_step` of the line is the un-instrumented public call",
            "instrumented_ms_per_step": instr_ms, "instrumented_clocks": clocks, "share_of_instrumented_step": (attn_ms / instr_ms) if instr_ms else None,
            "share_of_step": attn_ms / ms_per_step}


# ------------------------------------------------------------------------------------------------ plumbing
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
            time.sleep(0.002)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def bind_near_gpu(index):
    """Pin this process to the CPU cores NVML reports as local to GPU `index` (same NUMA node / PCIe root), BEFORE the
    pinned host buffers of the e2e leg are allocated, so that first-touch places them next to the GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        numa = None
        try:
            numa = pynvml.nvmlDeviceGetNumaNodeId(h)
        except Exception:
            pass
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"cpus": len(cpus), "numa_node": numa}
    except Exception:
        pass
    return None


def dist_setup():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier(world, dev):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize(dev)


def max_over_ranks(vals, world, dev):
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def gather_floats(val, world, dev):
    t = torch.zeros(world, dtype=torch.float64, device=dev)
    t[int(os.environ.get("RANK", "0"))] = val
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t)
    return t.tolist()


def api_fn(L, qk, pv):
    if qk == "mixed":
        return L.lowbit_fa_q_int8_k_dynamic
    if pv == "fp8":
        return L.lowbit_fa_qk_int4_pv_fp8 if qk != "int8" else L.lowbit_fa_qk_int8_pv_fp8_cuda
    return {"int8": L.lowbit_fa_qk_int8_pv_fp16_triton, "int4": L.lowbit_fa_qk_int4_pv_fp16_triton,
            "q8k4": L.lowbit_fa_q_int8_k_int4_pv_fp16}[qk]


def make_inputs(B, Hq, Hkv, N, D, layout, qk, dev, seed=0):
    torch.manual_seed(seed)
    shp = lambda h: (B, h, N, D) if layout == "HND" else (B, N, h, D)
    q, k, v = (torch.randn(shp(h), dtype=torch.float16, device=dev) for h in (Hq, Hkv, Hkv))
    if qk == "mixed":
        # dynamic-K workloads: every other 64-key block at 0.2x magnitude, so that the block statistic max|k|/127 puts
        # it in the INT2 class (<= 0.0125) and the rest in the INT4 class (randn: ~0.035)
        w = torch.where((torch.arange(N, device=dev) // 64) % 2 == 1, 0.2, 1.0).to(k.dtype)
        k = k * (w.view(1, 1, N, 1) if layout == "HND" else w.view(1, N, 1, 1))
    return q, k, v


def timed_steps(step, K, W, world, dev, local, with_clocks=True):
    """W warm-up steps, then K steps between barrier + synchronize on both sides, CUDA events on the launching stream;
    returns (ms per step on this rank, clock summary)."""
    stream = torch.cuda.current_stream(dev)
    for _ in range(W):
        step(None)
    sampler = ClockSampler(local) if with_clocks else None
    barrier(world, dev)
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(K):
        step(i)
    e1.record(stream)
    barrier(world, dev)
    if sampler:
        sampler.stop_flag = True
    return e0.elapsed_time(e1) / K, (sampler.summary() if sampler else None)


def decomposed_step(q, k, v, layout, causal, qk, K, dev):
    """The operator call cut at the attention launch so that CUDA events can bracket the dominant kernel inside the
    timed steps: == lowbit_fa_qk_int8_pv_fp16_triton / lowbit_fa_q_int8_k_int4_pv_fp16 (checked bit for bit)."""
    from lowbit_quant_fa2_paddle_b200 import _native as NV
    from lowbit_quant_fa2_paddle_b200 import attention as A
    from lowbit_quant_fa2_paddle_b200 import quant as Qz
    D = q.shape[-1]
    sm_scale = 1.0 / D ** 0.5
    kbits, packed = (8, False) if qk == "int8" else (4, True)
    qk_mode = NV.QK_Q8K4 if packed else NV.QK_I8
    stream = torch.cuda.current_stream(dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]

    def step(i=None):
        qc, qs, kc, ks, _ = Qz.smooth_and_quantize(q, k, True, sm_scale, layout, 8, kbits, packed, "triton")
        if i is not None:
            ev[i][0].record(stream)
        o, _ = A._forward(qc, kc, v, qs, ks, layout, torch.float16, False, causal, qk_mode=qk_mode)
        if i is not None:
            ev[i][1].record(stream)
        return o

    def attn_only(reps):
        qc, qs, kc, ks, _ = Qz.smooth_and_quantize(q, k, True, sm_scale, layout, 8, kbits, packed, "triton")
        f = lambda: A._forward(qc, kc, v, qs, ks, layout, torch.float16, False, causal, qk_mode=qk_mode)
        for _ in range(3):
            f()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(reps):
            f()
        a1.record(stream)
        torch.cuda.synchronize(dev)
        return a0.elapsed_time(a1) / reps
    return step, ev, attn_only


def e2e_measure(L, fn, q, k, v, layout, causal, KE, world, dev):
    """End to end from pinned host memory through lowbit_fa_host (H2D q,k,v + hot path + D2H o per step)."""
    stream = torch.cuda.current_stream(dev)
    hq_, hk_, hv_ = (t.contiguous().cpu().pin_memory() for t in (q, k, v))
    ho = torch.empty(hq_.shape, dtype=torch.float16).pin_memory()

    def serial():
        dq, dk, dv = (t.to(dev, non_blocking=True) for t in (hq_, hk_, hv_))
        ho.copy_(fn(dq, dk, dv, tensor_layout=layout, is_causal=causal), non_blocking=True)

    serial()
    torch.cuda.synchronize(dev)
    ho_serial = ho.clone()
    y0, y1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    y0.record(stream)
    for _ in range(3):
        serial()
    y1.record(stream)
    torch.cuda.synchronize(dev)
    serial_ms = y0.elapsed_time(y1) / 3
    ho.zero_()
    run = lambda: L.lowbit_fa_host(hq_, hk_, hv_, out=ho, op=fn, tensor_layout=layout, is_causal=causal, graph=True)
    for _ in range(3):
        run()
    barrier(world, dev)
    xe = [torch.cuda.Event(enable_timing=True) for _ in range(KE + 1)]
    xe[0].record(stream)
    for i in range(KE):
        run()
        xe[i + 1].record(stream)  # lowbit_fa_host leaves the caller's stream waiting for the last copy-out
    barrier(world, dev)
    per = sorted(xe[i].elapsed_time(xe[i + 1]) for i in range(KE))
    assert torch.equal(ho, ho_serial), "pipelined host entry point differs from the serial call"
    # plain copy rate of this rank's inputs with every rank copying at the same time
    dq = torch.empty(hq_.shape, dtype=hq_.dtype, device=dev)
    barrier(world, dev)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(stream)
    for _ in range(4):
        dq.copy_(hq_, non_blocking=True)
    c1.record(stream)
    barrier(world, dev)
    h2d_gbs = 4 * hq_.numel() * 2 / (c0.elapsed_time(c1) * 1e-3) / 1e9
    return {"median_ms": per[KE // 2], "mean_ms": xe[0].elapsed_time(xe[KE]) / KE, "max_ms": per[-1], "serial_ms": serial_ms,
            "steps": KE, "h2d": int(hq_.numel() * 2 + hk_.numel() * 2 + hv_.numel() * 2), "d2h": int(ho.numel() * 2),
            "h2d_gbs": h2d_gbs}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_units(wl, units, reps=1):
    """Time the oracle port (reference pure-Paddle math restated on torch-CPU) on `units` (batch, head) units of the
    workload with all host threads; returns (TOPS, cores, sample description, seconds per rep)."""
    from oracle import attention as OA
    B, Hq, Hkv, N, D, layout, causal, qk, pv, part, desc = WL[wl]
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    g = torch.Generator().manual_seed(0)
    shp = (1, units, N, D) if layout == "HND" else (1, N, units, D)
    q, k, v = (torch.randn(shp, generator=g).half() for _ in range(3))
    kbits = 4 if qk in ("q8k4", "int4") else 8
    grp = 4 if N <= 4096 else (2 if N <= 8192 else 1)  # bounds the fp32 score matrix to ~1-1.3 GiB
    t0 = time.perf_counter()
    for _ in range(reps):
        for h0 in range(0, units, grp):
            sl = (slice(None), slice(h0, h0 + grp)) if layout == "HND" else (slice(None), slice(None), slice(h0, h0 + grp))
            OA.cpu_quantize_and_attend(q[sl], k[sl], v[sl], layout, causal, kbits=kbits)
    dt = (time.perf_counter() - t0) / reps
    ops = 4.0 * units * N * N * D / (2 if causal else 1)
    total = B * Hq
    return ops / dt / 1e12, cores, (f"{units} of the {total} (batch, head) units of {wl} per step (N={N}, D={D}, fp32 math, "
                                    f"all {cores} host threads); TOPS is a rate, so it needs no scaling"), dt


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return  # the other ranks of a torchrun launch exit 0 without work
    wl = args.workload or ("c2" if args.gpus == 1 else "c4s")
    B, Hq, Hkv, N, D, layout, causal, qk, pv, part, desc = WL[wl]
    if pv != "fp16" or qk not in ("int8", "q8k4"):
        raise SystemExit("--impl reference is defined for the INT8-QK / q8k4, fp16-PV workloads")
    units = B * Hq if N <= 4096 else (8 if N <= 8192 else 2)  # config 2 in full; a bounded sample of the long ones
    for _ in range(args.warmup):
        cpu_units(wl, min(units, 4))
    vals = []
    for _ in range(args.steps):
        vals.append(cpu_units(wl, units))
    tops = sum(v[0] for v in vals) / len(vals)
    ms = sum(v[3] for v in vals) / len(vals) * 1e3
    cores, sample = vals[0][1], vals[0][2]
    line = {"impl": "reference", "metric": METRIC, "value": tops, "unit": "TOPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None,
            "dtype": f"{qk} QK / {pv} PV codes, fp32 math on the CPU", "data": "synthetic randn fp16 seed 0",
            "config": config_of(wl, args.gpus),
            "cpu_baseline": {"value": tops, "unit": "TOPS", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": tops, "unit": "TOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def config_of(wl, n_gpus):
    """The `config` object: identical for our arm and the reference arm at the same N."""
    B, Hq, Hkv, N, D, layout, causal, qk, pv, part, desc = WL[wl]
    part_txt = {"replicate": "every rank the whole workload" if n_gpus > 1 else "single GPU",
                "heads": f"kv heads split over {n_gpus} rank(s), no collective on the data path",
                "ring": f"sequence split over {n_gpus} rank(s), NCCL P2P ring of quantized K/V"}[part]
    return {"workload": desc, "B": B, "Hq": Hq, "Hkv": Hkv, "N": N, "D": D, "layout": layout, "causal": causal,
            "qk": qk, "pv": pv, "partition": part_txt, "smooth_k": True, "quantization_backend": "triton (Q1 rounding)",
            "l2": "working set larger than the 126 MB L2, no flush"}


# ------------------------------------------------------------------------------------------------ reference kernels on this GPU
def reference_gpu(q, k, v, causal, ops):
    """The reference's own Triton kernels (staged unmodified under oracle/_ref by oracle/stage_ref.sh), JIT-compiled for
    this GPU: attention-only and quantize + attention TOPS at the bench workload."""
    try:
        os.environ["TRITON_INTERPRET"] = "0"
        os.environ.setdefault("LOWBIT_REFERENCE_ROOT", os.path.join(ROOT, "oracle", "_ref", "reference"))
        from oracle import ref_triton as RT
        if not RT.available():
            return {"unavailable": "reference kernels not staged under oracle/_ref (oracle/stage_ref.sh runs in the build container)"}
        import lowbit_quant_fa2_paddle_b200 as L
        stream = torch.cuda.current_stream(q.device)

        def timed(f, reps=10, warm=3):
            for _ in range(warm):
                f()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                f()
            b.record(stream)
            torch.cuda.synchronize(q.device)
            return a.elapsed_time(b) / reps
        km = L.k_mean(k)
        qi, qs, ki, ks = RT.per_block_int8(q, k, km=km)
        t_attn = timed(lambda: RT.attn_forward(qi, ki, v, qs, ks, "HND", causal, torch.float16, False))

        def op():
            a, s1, c, s2 = RT.per_block_int8(q, k, km=k.mean(dim=2, keepdim=True))
            return RT.attn_forward(a, c, v, s1, s2, "HND", causal, torch.float16, False)[0]
        t_op = timed(op)
        ours = L.lowbit_fa_qk_int8_pv_fp16_triton(q, k, v, is_causal=causal)
        ref = op()
        cos = float(torch.nn.functional.cosine_similarity(ours.float().flatten(), ref.float().flatten(), dim=0))
        return {"kind": "reference Triton kernels (src/triton/quant_per_block.py, attn_qk_int8_per_block*.py) JIT-compiled for sm_100",
                "attn_only_tops": ops / t_attn / 1e9, "value_tops": ops / t_op / 1e9, "attn_ms": t_attn, "op_ms": t_op,
                "our_output_cos_vs_reference": cos,
                "our_output_max_abs_vs_reference": float((ours.float() - ref.float()).abs().max())}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


# ------------------------------------------------------------------------------------------------ our arm, N = 1
def quick_tops(L, wl, dev, reps=None):
    """One short timing of another BASELINE config through its public operator (no decomposition)."""
    B, Hq, Hkv, N, D, layout, causal, qk, pv, part, desc = WL[wl]
    q, k, v = make_inputs(B, Hq, Hkv, N, D, layout, qk, dev)
    fn = api_fn(L, qk, pv)
    f = lambda: fn(q, k, v, tensor_layout=layout, is_causal=causal)
    ops = 4.0 * B * Hq * N * N * D / (2 if causal else 1)
    reps = reps or (10 if ops < 3e12 else 3)
    stream = torch.cuda.current_stream(dev)
    for _ in range(3):
        f()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        f()
    b.record(stream)
    torch.cuda.synchronize(dev)
    ms = a.elapsed_time(b) / reps
    del q, k, v
    torch.cuda.empty_cache()
    return {"workload": desc, "value": ops / ms / 1e9, "unit": "TOPS", "ms_per_step": ms, "steps": reps}


def kv_decode_line(L, dev, reps=12):
    """one query row per (batch, head) against a KIVI-packed 4-bit K/V cache through
    quantized_flash_attn_forward; an HBM-bound path, so the figure is GB/s over the algorithmic bytes (codes + scales +
    minima, read once) against the measured copy bandwidth.  Three caches are rotated so that no launch finds its
    input in the 126 MB L2."""
    B, H, N, D, bits = 4, 32, 16384, 128, 4
    torch.manual_seed(0)
    caches = []
    for _ in range(3):
        k, v = (torch.randn(B, N, H, D, dtype=torch.float16, device=dev) for _ in range(2))
        caches.append(L.quant_and_pack_kv(k, v, 32, bits))
        del k, v
    q = torch.randn(B, 1, H, D, dtype=torch.float16, device=dev)
    nbytes = sum(t.numel() * t.element_size() for t in caches[0])
    f = lambda i: L.quantized_flash_attn_forward(q, *caches[i % 3], group_size=32, bits=bits)
    for i in range(3):
        f(i)
    stream = torch.cuda.current_stream(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for i in range(reps):
        f(i)
    b.record(stream)
    torch.cuda.synchronize(dev)
    ms = a.elapsed_time(b) / reps
    del caches, q
    torch.cuda.empty_cache()
    hbm = peaks()["hbm_gbs"]
    return {"workload": f"KV-cache decode: B{B} H{H} N{N} D{D}, {bits}-bit KIVI cache, 1 query row", "value": nbytes / ms / 1e6,
            "unit": "GB/s", "ms_per_step": ms, "steps": reps, "algorithmic_bytes": nbytes,
            "roofline": {"bound": "hbm", "achieved": nbytes / ms / 1e6, "peak": hbm, "unit": "GB/s", "frac": nbytes / ms / 1e6 / hbm}}


def run_single(args, rank, world, local, dev):
    """N = 1 (or --workload with partition "replicate" at N > 1: every rank the same work, weak scaling)."""
    affinity = bind_near_gpu(local)
    import lowbit_quant_fa2_paddle_b200 as L
    from lowbit_quant_fa2_paddle_b200 import _native
    _native.lib()
    wl = args.workload or "c2"
    B, Hq, Hkv, N, D, layout, causal, qk, pv, part, desc = WL[wl]
    W, K = max(args.warmup, 3), args.steps
    q, k, v = make_inputs(B, Hq, Hkv, N, D, layout, qk, dev, seed=rank)
    ops = 4.0 * B * Hq * N * N * D / (2 if causal else 1)
    fn = api_fn(L, qk, pv)
    decomposable = pv == "fp16" and qk in ("int8", "q8k4", "int4")
    attn_ev = attn_only = None
    if decomposable:
        step, attn_ev, attn_only = decomposed_step(q, k, v, layout, causal, qk, K, dev)
        assert torch.equal(fn(q, k, v, tensor_layout=layout, is_causal=causal), step()), "bench step differs from the public API call"
        launches = 5
    else:
        step = lambda i=None: fn(q, k, v, tensor_layout=layout, is_causal=causal)
        launches = None
    # headline: the public operator call (what a user makes: one C-ABI call per step), no events inside the steps
    api_step = lambda i=None: fn(q, k, v, tensor_layout=layout, is_causal=causal)
    ms_local, clocks = timed_steps(api_step, K, W, world, dev, local)
    # instrumented pass: the same K steps cut at the attention launch with a CUDA-event pair around it (the events and
    # the Python-side fork / join of the quantizer streams cost time, so this pass is not the headline)
    instr_ms_local, attn_ms = None, None
    instr_clocks = None
    if attn_ev:
        cool_down(dev)
        instr_ms_local, instr_clocks = timed_steps(step, K, W, world, dev, local)
        attn_ms = sum(a.elapsed_time(b) for a, b in attn_ev) / K
        cool_down(dev)
    attn_alone_ms = attn_only(max(10, min(K, 50))) if attn_only else None
    e2e = None
    if decomposable and wl in ("c2", "c2c", "c4s", "c4"):
        e2e = e2e_measure(L, fn, q, k, v, layout, causal, max(3, min(K, 20)), world, dev)
    ms_per_step, attn_ms_m, attn_alone_m, e2e_ms, instr_ms = max_over_ranks(
        [ms_local, attn_ms or 0.0, attn_alone_ms or 0.0, e2e["median_ms"] if e2e else 0.0, instr_ms_local or 0.0], world, dev)
    h2d_rates = gather_floats(e2e["h2d_gbs"], world, dev) if e2e else None
    if rank == 0:
        value = world * ops / (ms_per_step * 1e-3) / 1e12
        line = {"metric": METRIC, "value": value, "unit": "TOPS", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": value / BASELINE_MD_TOPS if (wl == "c2" and world == 1) else None,
                "dtype": f"{qk} QK (int32 acc) / {pv} PV (fp32 acc)", "data": "synthetic randn fp16 seed 0",
                "config": config_of(wl, world), "gpu_launches": (launches * K) if launches else None, "clocks": clocks}
        if wl == "c2":
            line["vs_baseline_note"] = "BASELINE.md 199.5 TFLOP/s is attention-kernel-only on unstated hardware; value includes quantization"
        if attn_alone_ms:
            line["attn_only"] = {"value": world * ops / (attn_alone_m * 1e-3) / 1e12, "unit": "TOPS", "ms": attn_alone_m}
            line["roofline"] = roofline("attn_fwd_n64_kernel" if D == 64 else "attn_fwd_kernel", B, Hq, Hkv, N, D, causal, qk, pv,
                                        attn_ms_m, (clocks or {}).get("sm_mhz"), ncu_traffic_bytes() if wl == "c2" else None)
            line["roofline"].update(instrumented(instr_ms, attn_ms_m, ms_per_step, instr_clocks))
        if e2e:
            line["e2e"] = {"value": world * ops / (e2e_ms * 1e-3) / 1e12, "unit": "TOPS", "ms": e2e_ms,
                           "api": "lowbit_fa_host(graph=True): pinned host q,k,v -> pinned host o; (batch, head-group) chunks on 3 streams, replayed as one CUDA graph",
                           "serial_ms": e2e["serial_ms"], "statistic": "median of per-step CUDA-event times, max over ranks",
                           "mean_ms": e2e["mean_ms"], "max_ms": e2e["max_ms"], "steps": e2e["steps"], "host_binding": affinity,
                           "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"], "h2d_gbs_per_rank": h2d_rates}
        if world == 1 and not args.workload:
            # the other single-GPU configs of BASELINE.json, one short timing each (same process, after the headline)
            line["other_configs"] = {n: quick_tops(L, n, dev) for n in ("c2c", "c3q_8k", "c4s")}
            try:
                line["other_configs"]["kv_decode"] = kv_decode_line(L, dev)
            except Exception as e:  # noqa: BLE001
                line["other_configs"]["kv_decode"] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.manual_seed(0)
            q0, k0, v0 = (torch.randn(B, Hq, N, D, dtype=torch.float16, device=dev) for _ in range(3))
            line["reference_gpu"] = reference_gpu(q0, k0, v0, causal, ops)
            if not args.no_cpu_baseline:
                tops, cores, sample, _ = cpu_units(wl, 32 if N <= 4096 else 4)
                line["cpu_baseline"] = {"value": tops, "unit": "TOPS", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm, N > 1
def run_ring(L, P, wl, world, rank, dev, steps=5, warm=2):
    B, Hq, Hkv, N, D, layout, causal, qk, pv, part, desc = WL[wl]
    n_loc = N // world
    torch.manual_seed(1000 + rank)
    shp = lambda h: (B, h, n_loc, D) if layout == "HND" else (B, n_loc, h, D)
    q, k, v = (torch.randn(shp(h), dtype=torch.float16, device=dev) for h in (Hq, Hkv, Hkv))
    if qk == "mixed":
        seq = 2 if layout == "HND" else 1
        w = torch.where((torch.arange(n_loc, device=dev) // 64) % 2 == 1, 0.2, 1.0).to(k.dtype)
        k = k * (w.view(1, 1, n_loc, 1) if seq == 2 else w.view(1, n_loc, 1, 1))
    ring_qk = qk if qk in ("int8", "mixed") else "int4"
    f = lambda t=None: P.ring_attention(q, k, v, tensor_layout=layout, is_causal=causal, qk=ring_qk, pv=pv, timings=t)
    for _ in range(warm):
        f()
    stream = torch.cuda.current_stream(dev)
    barrier(world, dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        f()
    e1.record(stream)
    barrier(world, dev)
    ms = max_over_ranks([e0.elapsed_time(e1) / steps], world, dev)[0]
    tm = {}
    f(tm)  # one more pass with phase events (it synchronises the device: outside the timed region)
    comp, expo = sum(tm["compute"]), sum(tm["p2p_exposed"])
    parts = max_over_ranks([tm["k_mean"], tm["quantize"], comp, expo, tm["finalize"]], world, dev)
    ops = 4.0 * B * Hq * N * N * D / 2
    del q, k, v
    torch.cuda.empty_cache()
    total = sum(parts)
    limiter = max(zip(parts, ("K mean (sum + all-reduce)", "quantize", "attention of the resident shards",
                              "P2P wait beyond the step's compute", "finalize")))[1]
    return {"workload": desc, "value": ops / ms / 1e9, "unit": "TOPS", "ms_per_step": ms, "steps": steps, "scaling": "strong",
            "phases_ms_max_over_ranks": {"k_mean_allreduce": parts[0], "quantize_q_k_v": parts[1], "attention_compute": parts[2],
                                         "p2p_exposed": parts[3], "finalize": parts[4], "sum": total},
            "per_step_ms_rank0": {"compute": tm["compute"], "p2p_exposed": tm["p2p_exposed"]},
            "p2p_bytes_sent_per_rank_per_step": tm["p2p_bytes_per_step"], "ring_steps": world, "limiter": limiter}


def run_multi(args, rank, world, local, dev):
    """N > 1: BASELINE config 4, kv heads split over the ranks (no collective), strong scaling; config 5 ring as extras."""
    affinity = bind_near_gpu(local)
    import lowbit_quant_fa2_paddle_b200 as L
    from lowbit_quant_fa2_paddle_b200 import _native
    from lowbit_quant_fa2_paddle_b200 import parallel as P
    _native.lib()
    wl = args.workload or "c4s"
    B, Hq, Hkv, N, D, layout, causal, qk, pv, part, desc = WL[wl]
    W, K = max(args.warmup, 3), args.steps
    ops_total = 4.0 * B * Hq * N * N * D / (2 if causal else 1)
    fn = api_fn(L, qk, pv)
    if part == "ring":
        r = run_ring(L, P, wl, world, rank, dev, steps=max(3, min(K, 10)), warm=W)
        if rank == 0:
            line = {"metric": METRIC, "value": r["value"], "unit": "TOPS", "n_gpus": world, "steps": r["steps"], "warmup": W,
                    "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": f"{qk} QK (int32 acc) / {pv} PV (fp32 acc)", "data": "synthetic randn fp16", "config": config_of(wl, world),
                    "ring": r, "gpu_launches": None}
            print(json.dumps(line), flush=True)
        return
    # every rank holds the full tensors (same seed) and works on its kv-head slice as a strided view
    q, k, v = make_inputs(B, Hq, Hkv, N, D, layout, qk, dev, seed=0)
    hdim = 1 if layout == "HND" else 2
    hq0, hq1, hkv0, hkv1 = P.head_shard(Hq, Hkv, world, rank)
    qs_, ks_, vs_ = q.narrow(hdim, hq0, hq1 - hq0), k.narrow(hdim, hkv0, hkv1 - hkv0), v.narrow(hdim, hkv0, hkv1 - hkv0)
    step, attn_ev, attn_only = decomposed_step(qs_, ks_, vs_, layout, causal, qk, K, dev)
    o_api = P.lowbit_fa_head_sharded(q, k, v, fn, world, rank, tensor_layout=layout, is_causal=causal)
    assert torch.equal(o_api, step()), "bench step differs from parallel.lowbit_fa_head_sharded"
    del o_api
    api_step = lambda i=None: P.lowbit_fa_head_sharded(q, k, v, fn, world, rank, tensor_layout=layout, is_causal=causal)
    ms_local, clocks = timed_steps(api_step, K, W, world, dev, local)            # headline: the public call
    cool_down(dev)
    instr_ms_local, instr_clocks = timed_steps(step, K, W, world, dev, local)   # instrumented pass (see run_single)
    attn_ms = sum(a.elapsed_time(b) for a, b in attn_ev) / K
    cool_down(dev)
    attn_alone_ms = attn_only(max(10, min(K, 30)))
    # the same workload on ONE GPU, in this run (rank 0; the other ranks wait at the barrier)
    single_ms = 0.0
    cool_down(dev)
    if rank == 0:
        fs = lambda i=None: fn(q, k, v, tensor_layout=layout, is_causal=causal)
        for _ in range(3):
            fs()
        stream = torch.cuda.current_stream(dev)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K1 = max(3, min(K, 10))
        a0.record(stream)
        for _ in range(K1):
            fs()
        a1.record(stream)
        torch.cuda.synchronize(dev)
        single_ms = a0.elapsed_time(a1) / K1
    barrier(world, dev)
    e2e = e2e_measure(L, fn, qs_, ks_, vs_, layout, causal, max(3, min(K, 20)), world, dev)
    ms_per_step, attn_ms_m, attn_alone_m, e2e_ms, single_ms, instr_ms = max_over_ranks(
        [ms_local, attn_ms, attn_alone_ms, e2e["median_ms"], single_ms, instr_ms_local], world, dev)
    h2d_rates = gather_floats(e2e["h2d_gbs"], world, dev)
    h2d_tot, d2h_tot = (int(x) for x in max_over_ranks([e2e["h2d"] * 1.0, e2e["d2h"] * 1.0], world, dev))
    del q, k, v, qs_, ks_, vs_
    torch.cuda.empty_cache()
    ring = {}
    for name in ("c5", "c5dyn"):
        try:
            ring[name] = run_ring(L, P, name, world, rank, dev)
        except Exception as e:  # noqa: BLE001
            ring[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if rank == 0:
        value = ops_total / (ms_per_step * 1e-3) / 1e12
        hq_loc = (hq1 - hq0)
        line = {"metric": METRIC, "value": value, "unit": "TOPS", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": f"{qk} QK (int32 acc) / {pv} PV (fp32 acc)", "data": "synthetic randn fp16 seed 0",
                "config": config_of(wl, world), "gpu_launches": 5 * K, "clocks": clocks,
                "strong_scaling": {"single_gpu": {"value": ops_total / (single_ms * 1e-3) / 1e12, "unit": "TOPS", "ms_per_step": single_ms,
                                                  "how": "the whole workload on rank 0's GPU, same run, same build"},
                                   "speedup_vs_single_gpu": single_ms / ms_per_step, "ranks": world},
                "attn_only": {"value": ops_total / (attn_alone_m * 1e-3) / 1e12, "unit": "TOPS", "ms": attn_alone_m},
                "e2e": {"value": ops_total / (e2e_ms * 1e-3) / 1e12, "unit": "TOPS", "ms": e2e_ms,
                        "api": "lowbit_fa_host(graph=True) per rank on its kv-head slice: pinned host q,k,v -> pinned host o",
                        "statistic": "median of per-step CUDA-event times, max over ranks", "steps": e2e["steps"],
                        "host_binding": affinity, "h2d_bytes_per_step": h2d_tot * world, "d2h_bytes_per_step": d2h_tot * world,
                        "h2d_gbs_per_rank": h2d_rates,
                        "note": "all ranks copy from one host memory system at once: the per-rank H2D rate above is what bounds e2e"},
                "roofline": roofline("attn_fwd_n64_kernel", B, hq_loc, hkv1 - hkv0, N, D, causal, qk, pv, attn_ms_m,
                                     (clocks or {}).get("sm_mhz")),
                "ring": ring}
        line["roofline"].update(instrumented(instr_ms, attn_ms_m, ms_per_step, instr_clocks))
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WL))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    rank, world, local = dist_setup()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    part = WL[args.workload][9] if args.workload else ("replicate" if world == 1 else "heads")
    if world > 1 and part != "replicate":
        run_multi(args, rank, world, local, dev)
    else:
        run_single(args, rank, world, local, dev)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
