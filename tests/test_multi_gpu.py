"""Multi-GPU parity (`-m gpu`; skipped on a box with one GPU).

* two ranks, one process per GPU over NCCL: the sequence-parallel ring of quantized K/V (every K format, fp16 and FP8
  P.V, causal zig-zag and non-causal) against the single-GPU operator AND against fp32 SDPA, and head sharding bit for
  bit against the single-GPU result -- tools/ring_check.py, which exits non-zero on any mismatch;
* one process driving two GPUs: every kernel family on cuda:0 and then on cuda:1 (the > 48 KB shared-memory opt-in is
  a per-device function attribute: a per-process "configured once" flag fails the first launch on the second GPU).
"""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _need_two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")


def test_ring_and_head_sharding_two_ranks():
    _need_two_gpus()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "ring_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith(("ring ", "head-sharded"))]
    assert r.returncode == 0, "\n".join(lines) + "\n" + r.stderr[-3000:]
    assert len(lines) >= 7 and all("MISMATCH" not in ln for ln in lines), "\n".join(lines)


def test_one_process_two_devices():
    _need_two_gpus()
    import lowbit_quant_fa2_paddle_b200 as L
    from oracle import attention as OA
    torch.manual_seed(3)
    for d, n in ((64, 512), (128, 384)):
        q, k, v = (torch.randn(1, 2, n, d).half() for _ in range(3))
        ref = OA.lowbit_fa_api(q, k, v, "HND", False, compat_tail=False, pv_accum="fp32")
        outs = []
        for idx in (0, 1, 0):
            dev = torch.device("cuda", idx)
            qd, kd, vd = q.to(dev), k.to(dev), v.to(dev)
            o = L.lowbit_fa_qk_int8_pv_fp16_triton(qd, kd, vd)                      # TMA quantizer + attention
            o4 = L.lowbit_fa_qk_int4_pv_fp16_triton(qd, kd, vd, is_causal=True)     # packed-K expander path
            o8 = L.lowbit_fa_qk_int8_pv_fp8_cuda(qd, kd, vd)                        # V -> e4m3 + FP8 P.V
            kc, ks, km = L.k_smooth_quant(kd) if L.k_smooth_quant_supported(kd) else (None, None, None)  # cluster kernel
            torch.cuda.synchronize(dev)
            assert (o.cpu().float() - ref.float()).abs().max().item() <= 4e-3
            assert not torch.isnan(o4).any() and not torch.isnan(o8).any()
            outs.append(o.cpu())
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), "results differ between devices"
