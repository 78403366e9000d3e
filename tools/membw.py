#!/usr/bin/env python
"""HBM bandwidth probes: pure write (fill), pure read (sum), copy — to place write-heavy kernels on the right roofline."""
import torch
n = 1 << 30  # 1 Gi bf16 = 2 GiB
a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
b = torch.empty(n, dtype=torch.bfloat16, device="cuda")


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


s = t(lambda: a.fill_(1.0)); print(f"pure write (fill_ 2 GiB):      {2 * n / s / 1e12:.2f} TB/s")
s = t(lambda: a.zero_()); print(f"pure write (memset 2 GiB):     {2 * n / s / 1e12:.2f} TB/s")
s = t(lambda: a.view(torch.int16).max()); print(f"pure read (max over 2 GiB):    {2 * n / s / 1e12:.2f} TB/s")
s = t(lambda: b.copy_(a)); print(f"copy (2 GiB read + 2 GiB write): {4 * n / s / 1e12:.2f} TB/s")
m = 68664 * 2048
c = torch.empty(m, dtype=torch.bfloat16, device="cuda")
s = t(lambda: c.fill_(1.0), 50); print(f"pure write (fill_ 281 MB):     {2 * m / s / 1e12:.2f} TB/s  ({s * 1e6:.1f} us)")
