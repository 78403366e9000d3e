#!/usr/bin/env python
"""Does chunking the FFN over token ranges keep the hidden activations in L2 between the producing and consuming GEMMs?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chadavit_b200 import ops
bf16 = torch.bfloat16
T, D, F, dev = 68664 * 2, 192, 2048, "cuda"
r = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(bf16)
y, w1, w2 = r(T, D), r(F, D), r(D, F)
y32 = torch.randn(T, D, device=dev)
b1, b2 = torch.randn(F, device=dev), torch.randn(D, device=dev)
hid = torch.empty(T, F, device=dev, dtype=bf16)
z2 = torch.empty(T, D, device=dev)
dz2h, dh = r(T, D), torch.empty(T, F, device=dev, dtype=bf16)
dz2 = torch.randn(T, D, device=dev)
dy = torch.empty(T, D, device=dev)
gw1, gw2, gb1 = torch.zeros(F, D, device=dev), torch.zeros(D, F, device=dev), torch.zeros(F, device=dev)
R = ops.EPI_RESIDUAL_F32 | ops.EPI_OUT_F32


def fwd(nchunk):
    step = (T // nchunk + 127) // 128 * 128
    for c0 in range(0, T, step):
        c1 = min(T, c0 + step)
        ops.gemm(y[c0:c1], w1, bias=b1, flags=ops.EPI_RELU, out=hid[c0:c1])
        ops.gemm(hid[c0:c1], w2, bias=b2, aux=y32[c0:c1], flags=R, out=z2[c0:c1])


def bwd(nchunk):
    step = (T // nchunk + 127) // 128 * 128
    for c0 in range(0, T, step):
        c1 = min(T, c0 + step)
        n = c1 - c0
        ops.gemm(dz2h[c0:c1], hid[c0:c1], a_mn=True, b_mn=True, flags=ops.EPI_ATOMIC, out=gw2, k_splits=ops.splitk_for(n, 2 * 8))
        ops.gemm(dz2h[c0:c1], w2, b_mn=True, aux=hid[c0:c1], flags=ops.EPI_RELU_MASK, colsum=gb1, out=dh[c0:c1])
        ops.gemm(dh[c0:c1], y[c0:c1], a_mn=True, b_mn=True, flags=ops.EPI_ATOMIC, out=gw1, k_splits=ops.splitk_for(n, 16 * 2))
        ops.gemm(dh[c0:c1], w1, b_mn=True, aux=dz2[c0:c1], flags=R, out=dy[c0:c1])


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
    return e0.elapsed_time(e1) / reps * 1e3


for n in (1, 2, 4, 8, 16, 32):
    print(f"chunks={n:2d}  FFN fwd (fc1+fc2) {t(lambda: fwd(n)):8.1f} us   FFN bwd (dW2, dh, dW1, dy) {t(lambda: bwd(n)):8.1f} us   [T={T}]")
