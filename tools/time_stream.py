#!/usr/bin/env python
"""What a 67 MB (config-2 sized) streaming kernel can reach on this GPU: torch's own reduction / copy kernels over
rotating fp16 tensors (no launch finds its input in L2), next to the quantizer timings of tools/time_quant.py."""
import torch

dev = torch.device("cuda:0")
shape = (4, 32, 4096, 64)
xs = [torch.randn(shape, dtype=torch.float16, device=dev) for _ in range(4)]
outs8 = [torch.empty(shape, dtype=torch.int8, device=dev) for _ in range(4)]
outs16 = [torch.empty(shape, dtype=torch.float16, device=dev) for _ in range(4)]
elem = xs[0].numel()


def timeit(fn, reps=40):
    for i in range(4):
        fn(i % 4)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(reps):
        fn(i % 4)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


cases = [
    ("sum over seq (2 B/elem read)", lambda i: xs[i].sum(dim=2), 2.0),
    ("amax(abs) full (2 B/elem read)", lambda i: xs[i].abs_().amax() if False else torch.amax(xs[i]), 2.0),
    ("fp16 -> int8 cast copy (3 B/elem)", lambda i: outs8[i].copy_(xs[i]), 3.0),
    ("fp16 copy (4 B/elem)", lambda i: outs16[i].copy_(xs[i]), 4.0),
    ("mul_ in place (4 B/elem)", lambda i: xs[i].mul_(1.0), 4.0),
]
for label, fn, bpe in cases:
    t = timeit(fn)
    print(f"{label:36s} {t * 1e6:7.1f} us  {elem * bpe / t / 1e9:7.0f} GB/s", flush=True)
