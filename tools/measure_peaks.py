#!/usr/bin/env python
"""tools/measure_peaks.py -- dense INT8 / FP8 / FP16 tensor-core peaks of THIS B200, measured the way the driver
measured MEASURED_PEAKS.json's bf16 figure (library GEMM 8192^3, 2*N^3 ops; best of 10 launches = burst, back-to-back
for ~3 s = sustained under the power cap, SM clock sampled).  Writes profiles/peaks_int8_fp8.json -- the denominators
of bench.py's `roofline.t_min` (QK at the INT8 rate, PV at the FP16 / FP8 rate).

    python tools/measure_peaks.py [out.json]
"""
import json
import os
import sys
import threading
import time

import torch

N = 8192


def clock_sampler(stop, out):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        out["sm_max_mhz"] = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        while not stop.is_set():
            out.setdefault("sm", []).append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            out.setdefault("w", []).append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
            time.sleep(0.02)
    except Exception as e:  # pragma: no cover
        out["error"] = repr(e)


def measure(name, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    ops = 2.0 * N ** 3
    stop, clk = threading.Event(), {}
    th = threading.Thread(target=clock_sampler, args=(stop, clk), daemon=True)
    th.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(10, int(3000.0 / best))
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    stop.set()
    th.join()
    sus = a.elapsed_time(b) / reps
    sm = sorted(clk.get("sm", []))
    r = {"burst_tops": ops / best / 1e9, "sustained_tops": ops / sus / 1e9, "burst_ms": best, "sustained_ms": sus,
         "reps": reps, "sm_mhz_median_sustained": sm[len(sm) // 2] if sm else None,
         "sm_max_mhz": clk.get("sm_max_mhz"), "power_w_max": max(clk.get("w", [0]))}
    print(name, json.dumps(r), flush=True)
    return r


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "n": N,
           "how": "library GEMM 8192^3 (2*N^3 ops): best of 10 launches (burst) and back to back for ~3 s (sustained); "
                  "int8: torch._int_mm (cuBLASLt s8*s8->s32); fp8: torch._scaled_mm e4m3*e4m3->bf16, unit scales; "
                  "fp16/bf16: torch.matmul", "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    a8 = torch.randint(-127, 128, (N, N), dtype=torch.int8, device=dev)
    b8 = torch.randint(-127, 128, (N, N), dtype=torch.int8, device=dev)
    try:
        out["int8"] = measure("int8", lambda: torch._int_mm(a8, b8))
    except Exception as e:
        out["int8"] = {"error": repr(e)}
    try:
        af = torch.randn(N, N, device=dev).to(torch.float8_e4m3fn)
        bf = torch.randn(N, N, device=dev).to(torch.float8_e4m3fn).t()  # column-major second operand
        one = torch.ones((), device=dev)
        out["fp8_e4m3"] = measure("fp8", lambda: torch._scaled_mm(af, bf, scale_a=one, scale_b=one, out_dtype=torch.bfloat16))
    except Exception as e:
        out["fp8_e4m3"] = {"error": repr(e)}
    for name, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
        x = torch.randn(N, N, device=dev, dtype=dt)
        y = torch.randn(N, N, device=dev, dtype=dt)
        out[name] = measure(name, lambda: torch.matmul(x, y))
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                                                                "profiles", "peaks_int8_fp8.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
