#!/usr/bin/env python
"""PCIe floor of the e2e pipeline at config 2 (development aid): the same 9 x (K, Q, V) pinned-host -> device copies that
lowbit_fa_host issues, alone and with the 9 device -> host copies of O running on a second stream; no kernels."""
import torch

dev = torch.device("cuda:0")
B, H, N, D = 4, 32, 4096, 64
hq, hk, hv = (torch.randn(B, H, N, D, dtype=torch.float16).pin_memory() for _ in range(3))
ho = torch.empty(B, H, N, D, dtype=torch.float16).pin_memory()
dq, dk, dv, do = (torch.empty(B, H, N, D, dtype=torch.float16, device=dev) for _ in range(4))
plan = [(b, h0, h0 + 16) for b in range(B) for h0 in (0, 16)]
s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def run(with_d2h, lag=1):
    cur = torch.cuda.current_stream(dev)
    s_in.wait_stream(cur)
    s_out.wait_stream(cur)
    evs = []
    for i, (b, h0, h1) in enumerate(plan):
        with torch.cuda.stream(s_in):
            for d, h in ((dk, hk), (dq, hq), (dv, hv)):
                d[b, h0:h1].copy_(h[b, h0:h1], non_blocking=True)
            e = torch.cuda.Event()
            e.record(s_in)
            evs.append(e)
        if with_d2h:
            with torch.cuda.stream(s_out):
                s_out.wait_event(evs[i])
                ho[b, h0:h1].copy_(do[b, h0:h1], non_blocking=True)
    cur.wait_stream(s_in)
    cur.wait_stream(s_out)


for name, w in (("H2D only (201 MB in 24 copies)", False), ("H2D + concurrent D2H (67 MB in 8 copies)", True)):
    for _ in range(3):
        run(w)
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(10):
        run(w)
    z.record()
    torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(z) / 10:.3f} ms per pass", flush=True)
