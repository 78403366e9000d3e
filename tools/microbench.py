#!/usr/bin/env python
"""Per-kernel microbenchmarks at the shapes of one moyen/16 global-crop pass (T ~ 68.7k tokens, D=192, F=2048).

    python tools/microbench.py [--tokens 68664]

Each op is timed alone with CUDA events (10 warm-ups, 30 reps, buffers >> L2 are rotated for HBM-bound ops);
prints us, TFLOP/s and algorithmic GB/s so the GEMM/attention/row-wise kernels can be compared with the roofline."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chadavit_b200 import ops  # noqa: E402

bf16 = torch.bfloat16


def timeit(fn, reps=30, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def report(name, us, flops=0.0, bytes_=0.0):
    print(f"{name:46s} {us:9.1f} us  {flops / us / 1e6:8.1f} TFLOP/s  {bytes_ / us / 1e3:8.1f} GB/s")


def optim_section(dev):
    """LARS path of the fused step at the size of backbone + head (17.5 M parameters): norms (8 B/param) + step (40 B/param)."""
    n = 17_510_464
    p, gr, buf, t = (torch.randn(n, device=dev) for _ in range(4))
    p16, t16 = torch.empty(n, device=dev, dtype=bf16), torch.empty(n, device=dev, dtype=bf16)
    nseg = 160
    start = torch.linspace(0, n // 64, nseg + 1).to(torch.int32).to(dev)
    seg_of = torch.repeat_interleave(torch.arange(nseg, dtype=torch.int32), (start[1:] - start[:-1]).long().cpu()).to(dev)
    flags = torch.full((n,), 5, dtype=torch.uint8, device=dev)
    partial, norms = torch.empty(n // 32, device=dev), torch.empty(3 * nseg, device=dev)
    clipf = torch.ones(nseg, dtype=torch.uint8, device=dev)
    report("param_norms 17.5M params (8 B/param), 160 segs", timeit(lambda: ops.param_norms(p, gr, start, clipf, partial, norms, clip=3.0)), 0, n * 8.0)
    report("lars+ema+bf16 shadows 17.5M params (40 B/param)",
           timeit(lambda: ops.lars_step(p, gr, buf, flags, seg_of, norms, lr=1e-3, momentum=0.9, weight_decay=1e-6, eta=0.02, clip_lr=True,
                                        p_bf16=p16, teacher=t, teacher_bf16=t16, tau=0.99)), 0, n * 40.0)
    report("scale_grads 17.5M params (8 B/param)", timeit(lambda: ops.scale_grads(gr, seg_of, norms)), 0, n * 8.0)


def ffn_section(dev, T):
    D, F = 192, 2048
    r = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(bf16)  # noqa: E731
    x16, w1, w2 = r(T, D), r(F, D), r(D, F)
    x32 = torch.randn(T, D, device=dev)
    b1, b2 = torch.randn(F, device=dev), torch.randn(D, device=dev)
    for TT in (T, 2 * T):
        xa, xr = (x16, x32) if TT == T else (torch.cat([x16, x16]), torch.cat([x32, x32]))
        report(f"ffn fused T={TT} (no hidden store)", timeit(lambda: ops.ffn_fwd(xa, w1, b1, w2, b2, xr, save_hidden=False)), 4.0 * TT * D * F, TT * (D * 2 + D * 8))
        report(f"ffn fused T={TT} (+ hidden store)", timeit(lambda: ops.ffn_fwd(xa, w1, b1, w2, b2, xr, save_hidden=True)), 4.0 * TT * D * F, TT * (D * 2 + D * 8 + F * 2))


def dh_section(dev, T):
    """d(hidden) = (dz2 W2) o (hidden > 0): mask read from the bf16 hidden activations vs from the 1-bit mask of cb_ffn_fwd."""
    D, F = 192, 2048
    r = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(bf16)  # noqa: E731
    for TT in (T, 2 * T):
        dz, w2, hid = r(TT, D), r(D, F), torch.relu(r(TT, F))
        bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (F // 32, (TT + 31) // 32 * 32), device=dev, dtype=torch.int32)
        o, cs = torch.empty(TT, F, device=dev, dtype=bf16), torch.zeros(F, device=dev)
        report(f"dh T={TT} mask = bf16 hidden", timeit(lambda: ops.gemm(dz, w2, b_mn=True, aux=hid, flags=ops.EPI_RELU_MASK, out=o, colsum=cs)),
               2.0 * TT * D * F, TT * (D + 2 * F) * 2)
        report(f"dh T={TT} mask = bits", timeit(lambda: ops.gemm(dz, w2, b_mn=True, aux=bits, flags=ops.EPI_RELU_MASK | ops.EPI_MASK_BITS, out=o, colsum=cs)),
               2.0 * TT * D * F, TT * (D * 2 + F * 2 + F / 8))


def dw_section(dev, T):
    """Bias gradients next to the weight gradients: d(hidden) with / without the column sums in its epilogue, the dW1 / dWin
    split-K products with / without the tensor-pipe sums of their A operand, and the separate column-sum pass they replace."""
    D, F = 192, 2048
    r = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(bf16)  # noqa: E731
    for TT in (T, 2 * T):
        dz, w2, y, dh, dqkv = r(TT, D), r(D, F), r(TT, D), r(TT, F), r(TT, 3 * D)
        bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (F // 32, (TT + 31) // 32 * 32), device=dev, dtype=torch.int32)
        o, cs = torch.empty(TT, F, device=dev, dtype=bf16), torch.zeros(F, device=dev)
        M = ops.EPI_RELU_MASK | ops.EPI_MASK_BITS
        report(f"dh T={TT} bits, colsum in the epilogue", timeit(lambda: ops.gemm(dz, w2, b_mn=True, aux=bits, flags=M, out=o, colsum=cs)), 2.0 * TT * D * F, TT * (D * 2 + F * 2 + F / 8))
        report(f"dh T={TT} bits, no colsum", timeit(lambda: ops.gemm(dz, w2, b_mn=True, aux=bits, flags=M, out=o)), 2.0 * TT * D * F, TT * (D * 2 + F * 2 + F / 8))
        g1, gq, cq = torch.zeros(F, D, device=dev), torch.zeros(3 * D, D, device=dev), torch.zeros(3 * D, device=dev)
        k1, kq = ops.splitk_for(TT, 16 * 2), ops.splitk_for(TT, 5 * 2)
        A = ops.EPI_ATOMIC
        report(f"dW1 T={TT} split-K {k1}", timeit(lambda: ops.gemm(dh, y, a_mn=True, b_mn=True, flags=A, out=g1, k_splits=k1)), 2.0 * TT * D * F, TT * (D + F) * 2)
        report(f"dW1 T={TT} split-K {k1} + db1 (N = 208)", timeit(lambda: ops.gemm(dh, y, a_mn=True, b_mn=True, flags=A, out=g1, k_splits=k1, colsum=cs)), 2.0 * TT * D * F, TT * (D + F) * 2)
        report(f"dWin T={TT} split-K {kq}", timeit(lambda: ops.gemm(dqkv, y, a_mn=True, b_mn=True, flags=A, out=gq, k_splits=kq)), 2.0 * TT * D * 3 * D, TT * 4 * D * 2)
        report(f"dWin T={TT} split-K {kq} + db_in (N = 208)", timeit(lambda: ops.gemm(dqkv, y, a_mn=True, b_mn=True, flags=A, out=gq, k_splits=kq, colsum=cq)), 2.0 * TT * D * 3 * D, TT * 4 * D * 2)
        report(f"colsum bf16 [T={TT},576] (replaced)", timeit(lambda: ops.colsum(dqkv, cq)), 0, TT * 3 * D * 2)


def ffnbwd_section(dev, T):
    """Backward through the hidden layer: the fused kernel against the two products it replaces (same operands)."""
    D, F = 192, 2048
    r = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(bf16)  # noqa: E731
    for TT in (T, 2 * T):
        dz, w2, w1 = r(TT, D), r(D, F), r(F, D)
        dz32 = torch.randn(TT, D, device=dev)
        bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (F // 32, (TT + 31) // 32 * 32), device=dev, dtype=torch.int32)
        o, dy = torch.empty(TT, F, device=dev, dtype=bf16), torch.empty(TT, D, device=dev)
        M, R = ops.EPI_RELU_MASK | ops.EPI_MASK_BITS, ops.EPI_RESIDUAL_F32 | ops.EPI_OUT_F32
        t1 = timeit(lambda: ops.gemm(dz, w2, b_mn=True, aux=bits, flags=M, out=o))
        t2 = timeit(lambda: ops.gemm(o, w1, b_mn=True, aux=dz32, flags=R, out=dy))
        report(f"dh T={TT} (bits, no colsum)", t1, 2.0 * TT * D * F, TT * (D * 2 + F * 2 + F / 8))
        report(f"dy T={TT} [T,2048]x[2048,192] +res32", t2, 2.0 * TT * D * F, TT * (F * 2 + D * 8))
        report(f"  sum of the two T={TT}", t1 + t2, 4.0 * TT * D * F, TT * (D * 10 + F * 4 + F / 8))
        report(f"ffn bwd fused T={TT} (dh stored once)", timeit(lambda: ops.ffn_bwd(dz, w2, w1, bits, dz32)), 4.0 * TT * D * F, TT * (D * 10 + F * 2 + F / 8))
        b1, b2 = torch.randn(F, device=dev), torch.randn(D, device=dev)
        report(f"ffn fwd + hidden store + bits T={TT} (for scale)", timeit(lambda: ops.ffn_fwd(dz, w1, b1, w2, b2, dz32, save_hidden=True, save_mask_bits=True)), 4.0 * TT * D * F)


def kmajor_section(dev, T):
    """Input-gradient products dX = dY W: the weight read in place as an MN-major B operand ([K, N] row-major) against a
    pre-transposed K-major copy ([N, K]) — same math, same bytes, only the shared-memory layout the tensor core reads differs."""
    D, F = 192, 2048
    r = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(bf16)  # noqa: E731
    dz, dhid, dqkv, x32 = r(T, D), r(T, F), r(T, 3 * D), torch.randn(T, D, device=dev)
    w2, w1, win, wo = r(D, F), r(F, D), r(3 * D, D), r(D, D)
    bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (F // 32, (T + 31) // 32 * 32), device=dev, dtype=torch.int32)
    M, R = ops.EPI_RELU_MASK | ops.EPI_MASK_BITS, ops.EPI_RESIDUAL_F32 | ops.EPI_OUT_F32
    o = torch.empty(T, F, device=dev, dtype=bf16)
    w2t, w1t, wint, wot = (w.t().contiguous() for w in (w2, w1, win, wo))
    report("dh   B = W2 MN-major", timeit(lambda: ops.gemm(dz, w2, b_mn=True, aux=bits, flags=M, out=o)), 2.0 * T * D * F)
    report("dh   B = W2^T K-major", timeit(lambda: ops.gemm(dz, w2t, aux=bits, flags=M, out=o)), 2.0 * T * D * F)
    report("fc1-like (no mask) B MN-major", timeit(lambda: ops.gemm(dz, w2, b_mn=True, out=o)), 2.0 * T * D * F)
    report("fc1-like (no mask) B K-major", timeit(lambda: ops.gemm(dz, w2t, out=o)), 2.0 * T * D * F)
    o = torch.empty(T, D, device=dev)
    report("dy   B = W1 MN-major", timeit(lambda: ops.gemm(dhid, w1, b_mn=True, aux=x32, flags=R, out=o)), 2.0 * T * D * F)
    report("dy   B = W1^T K-major", timeit(lambda: ops.gemm(dhid, w1t, aux=x32, flags=R, out=o)), 2.0 * T * D * F)
    report("du   B = Win MN-major", timeit(lambda: ops.gemm(dqkv, win, b_mn=True, flags=ops.EPI_OUT_F32, out=o)), 2.0 * T * D * 3 * D)
    report("du   B = Win^T K-major", timeit(lambda: ops.gemm(dqkv, wint, flags=ops.EPI_OUT_F32, out=o)), 2.0 * T * D * 3 * D)
    o = torch.empty(T, D, device=dev, dtype=bf16)
    report("datt B = Wo MN-major", timeit(lambda: ops.gemm(dz, wo, b_mn=True, out=o)), 2.0 * T * D * D)
    report("datt B = Wo^T K-major", timeit(lambda: ops.gemm(dz, wot, out=o)), 2.0 * T * D * D)
    b1, b2 = torch.randn(F, device=dev), torch.randn(D, device=dev)
    report("ffn fwd + hidden store (K-major weights)", timeit(lambda: ops.ffn_fwd(dz, w1, b1, w2, b2, x32, save_hidden=True, save_mask_bits=True)), 4.0 * T * D * F)
    report("ffn bwd fused (MN-major weights)", timeit(lambda: ops.ffn_bwd(dz, w2, w1, bits, x32)), 4.0 * T * D * F)


def splitk_section(dev, T):
    """Split-K factor of the weight-gradient products: CTAs in flight = row tiles x column tiles x splits (148 SMs)."""
    D, F = 192, 2048
    r = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(bf16)  # noqa: E731
    for TT in (T, 2 * T):
        y, dh, dqkv, dz = r(TT, D), r(TT, F), r(TT, 3 * D), r(TT, D)
        A = ops.EPI_ATOMIC
        g1, g2, gq, go = torch.zeros(F, D, device=dev), torch.zeros(D, F, device=dev), torch.zeros(3 * D, D, device=dev), torch.zeros(D, D, device=dev)
        cs, cq = torch.zeros(F, device=dev), torch.zeros(3 * D, device=dev)
        for ks in (5, 9, 18):
            report(f"dW1  T={TT} [2048,T]x[T,192] 16 tiles x {ks} splits (+db1)", timeit(lambda: ops.gemm(dh, y, a_mn=True, b_mn=True, flags=A, out=g1, k_splits=ks, colsum=cs)), 2.0 * TT * D * F, TT * (D + F) * 2)
        for ks in (10, 9, 18):
            report(f"dW2  T={TT} [192,T]x[T,2048] 16 tiles x {ks} splits", timeit(lambda: ops.gemm(dz, dh, a_mn=True, b_mn=True, flags=A, out=g2, k_splits=ks)), 2.0 * TT * D * F, TT * (D + F) * 2)
        for ks in (15, 29, 58):
            report(f"dWin T={TT} [576,T]x[T,192] 5 tiles x {ks} splits (+db)", timeit(lambda: ops.gemm(dqkv, y, a_mn=True, b_mn=True, flags=A, out=gq, k_splits=ks, colsum=cq)), 2.0 * TT * D * 3 * D, TT * 4 * D * 2)
        for ks in (37, 74, 148):
            report(f"dWo  T={TT} [192,T]x[T,192] 2 tiles x {ks} splits", timeit(lambda: ops.gemm(dz, y, a_mn=True, b_mn=True, flags=A, out=go, k_splits=ks)), 2.0 * TT * D * D, TT * 2 * D * 2)


def attn_section(dev):
    """Attention forward / backward alone: the bench's ragged global-crop batch, 32 uniform 1961-token sequences, and the
    base/16 stress shape (D = 768, 12 heads of 64)."""
    r = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(bf16)  # noqa: E731
    for tag, counts, D, H in (("ragged 64 (bench)", np.random.RandomState(1234).randint(1, 11, size=64).tolist(), 192, 2),
                              ("ragged 128 (2 crops)", np.random.RandomState(1234).randint(1, 11, size=64).tolist() * 2, 192, 2),
                              ("32 x 1961", [10] * 32, 192, 2), ("base/16 64 x 1961", [10] * 64, 768, 12), ("moyen h12 64 ragged", np.random.RandomState(1234).randint(1, 11, size=64).tolist(), 192, 12)):
        lay = ops.PackedLayout(counts, 196, dev)
        qkv, do = r(lay.T, 3 * D), r(lay.T, D)
        out, lse = ops.attn_fwd(qkv, lay, H)
        report(f"attn fwd {tag} T={lay.T} H={H} d={D // H}", timeit(lambda: ops.attn_fwd(qkv, lay, H)), 4.0 * D * lay.sum_sq, lay.T * 4 * D * 2)
        report(f"attn bwd {tag} T={lay.T} H={H} d={D // H}", timeit(lambda: ops.attn_bwd(do, qkv, out, lse, lay, H)), 10.0 * D * lay.sum_sq, lay.T * 9 * D * 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=68664)
    ap.add_argument("--only", default="", help="comma list of sections: optim")
    a = ap.parse_args()
    if a.only == "optim":
        return optim_section("cuda")
    if a.only == "ffn":
        return ffn_section("cuda", a.tokens)
    if a.only == "dh":
        return dh_section("cuda", a.tokens)
    if a.only == "ffnbwd":
        return ffnbwd_section("cuda", a.tokens)
    if a.only == "kmajor":
        return kmajor_section("cuda", a.tokens)
    if a.only == "splitk":
        return splitk_section("cuda", a.tokens)
    if a.only == "dw":
        return dw_section("cuda", a.tokens)
    if a.only == "attn":
        return attn_section("cuda")
    T, D, F = a.tokens, 192, 2048
    dev = "cuda"
    r = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(bf16)  # noqa: E731
    x16, w_in, w_o, w1, w2 = r(T, D), r(3 * D, D), r(D, D), r(F, D), r(D, F)
    hid, qkv = r(T, F), r(T, 3 * D)
    x32 = torch.randn(T, D, device=dev)
    b_in, b_o, b1, b2 = (torch.randn(n, device=dev) for n in (3 * D, D, F, D))
    R = ops.EPI_RESIDUAL_F32 | ops.EPI_OUT_F32
    print(f"# T={T} D={D} F={F}")
    # ---- forward GEMMs
    o = torch.empty(T, 3 * D, device=dev, dtype=bf16)
    report("qkv   [T,192]x[192,576] +bias -> bf16", timeit(lambda: ops.gemm(x16, w_in, bias=b_in, out=o)), 2.0 * T * D * 3 * D, T * (D + 3 * D) * 2)
    o = torch.empty(T, D, device=dev)
    report("proj  [T,192]x[192,192] +bias +res32 -> f32", timeit(lambda: ops.gemm(x16, w_o, bias=b_o, aux=x32, flags=R, out=o)), 2.0 * T * D * D, T * D * (2 + 4 + 4))
    gam_, bet_ = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    report("proj + LayerNorm, two kernels (z kept)", timeit(lambda: ops.layernorm_fwd(ops.gemm(x16, w_o, bias=b_o, aux=x32, flags=R), gam_, bet_, 1e-5, out_f32=True)), 2.0 * T * D * D, T * D * (2 + 4 + 4 + 4 + 2 + 4))
    report("proj + LayerNorm fused (z kept: student)", timeit(lambda: ops.gemm_ln_fwd(x16, w_o, b_o, x32, gam_, bet_, 1e-5, keep_z=True)), 2.0 * T * D * D, T * D * (2 + 4 + 4 + 2 + 4))
    report("proj + LayerNorm fused (z dropped: no-grad)", timeit(lambda: ops.gemm_ln_fwd(x16, w_o, b_o, x32, gam_, bet_, 1e-5, keep_z=False, save_stats=False)), 2.0 * T * D * D, T * D * (2 + 4 + 2 + 4))
    o = torch.empty(T, F, device=dev, dtype=bf16)
    report("fc1   [T,192]x[192,2048] +bias relu -> bf16", timeit(lambda: ops.gemm(x16, w1, bias=b1, flags=ops.EPI_RELU, out=o)), 2.0 * T * D * F, T * (D + F) * 2)
    o = torch.empty(T, D, device=dev)
    report("fc2   [T,2048]x[2048,192] +bias +res32 -> f32", timeit(lambda: ops.gemm(hid, w2, bias=b2, aux=x32, flags=R, out=o)), 2.0 * T * D * F, T * (F * 2 + D * 8))
    zz = torch.empty(T, D, device=dev)
    report("ffn fused fc1+relu+fc2+res (no hidden store)", timeit(lambda: ops.ffn_fwd(x16, w1, b1, w2, b2, x32, save_hidden=False)), 4.0 * T * D * F, T * (D * 2 + D * 8))
    report("ffn fused fc1+relu+fc2+res (+ hidden store)", timeit(lambda: ops.ffn_fwd(x16, w1, b1, w2, b2, x32, save_hidden=True)), 4.0 * T * D * F, T * (D * 2 + D * 8 + F * 2))
    # ---- backward GEMMs
    o = torch.empty(T, F, device=dev, dtype=bf16)
    report("dh    [T,192]x[192,2048] (W2 MN) relu-mask", timeit(lambda: ops.gemm(x16, w2, b_mn=True, aux=hid, flags=ops.EPI_RELU_MASK, out=o)), 2.0 * T * D * F, T * (D + 2 * F) * 2)
    o = torch.empty(T, D, device=dev)
    report("dy    [T,2048]x[2048,192] (W1 MN) +res32 -> f32", timeit(lambda: ops.gemm(hid, w1, b_mn=True, aux=x32, flags=R, out=o)), 2.0 * T * D * F, T * (F * 2 + D * 8))
    o = torch.empty(T, D, device=dev)
    report("du    [T,576]x[576,192] (Win MN) -> f32", timeit(lambda: ops.gemm(qkv, w_in, b_mn=True, flags=ops.EPI_OUT_F32, out=o)), 2.0 * T * D * 3 * D, T * (3 * D * 2 + D * 4))
    g = torch.zeros(D, F, device=dev)
    ks = ops.splitk_for(T, 2 * 8)
    report(f"dW2   [192,T]x[T,2048] split-K {ks} atomic", timeit(lambda: ops.gemm(x16, hid, a_mn=True, b_mn=True, flags=ops.EPI_ATOMIC, out=g, k_splits=ks)), 2.0 * T * D * F, T * (D + F) * 2)
    g = torch.zeros(F, D, device=dev)
    ks = ops.splitk_for(T, 16 * 2)
    report(f"dW1   [2048,T]x[T,192] split-K {ks} atomic", timeit(lambda: ops.gemm(hid, x16, a_mn=True, b_mn=True, flags=ops.EPI_ATOMIC, out=g, k_splits=ks)), 2.0 * T * D * F, T * (D + F) * 2)
    g = torch.zeros(3 * D, D, device=dev)
    ks = ops.splitk_for(T, 5 * 2)
    report(f"dWin  [576,T]x[T,192] split-K {ks} atomic", timeit(lambda: ops.gemm(qkv, x16, a_mn=True, b_mn=True, flags=ops.EPI_ATOMIC, out=g, k_splits=ks)), 2.0 * T * D * 3 * D, T * 4 * D * 2)
    # ---- attention (ragged U{1..10} batch of 64 at 224^2 like bench.py)
    counts = np.random.RandomState(1234).randint(1, 11, size=64).tolist()
    lay = ops.PackedLayout(counts, 196, dev)
    qkv2 = r(lay.T, 3 * D)
    do = r(lay.T, D)
    out, lse = ops.attn_fwd(qkv2, lay, 2)
    report(f"attn fwd  T={lay.T} H=2 d=96", timeit(lambda: ops.attn_fwd(qkv2, lay, 2)), 4.0 * D * lay.sum_sq, lay.T * 4 * D * 2)
    report(f"attn bwd  T={lay.T} H=2 d=96", timeit(lambda: ops.attn_bwd(do, qkv2, out, lse, lay, 2)), 10.0 * D * lay.sum_sq, lay.T * 9 * D * 2)
    lay10 = ops.PackedLayout([10] * 32, 196, dev)
    qkv3 = r(lay10.T, 3 * D)
    do3 = r(lay10.T, D)
    out3, lse3 = ops.attn_fwd(qkv3, lay10, 2)
    report(f"attn fwd  32 x 1961 tokens H=2 d=96", timeit(lambda: ops.attn_fwd(qkv3, lay10, 2)), 4.0 * D * lay10.sum_sq)
    report(f"attn bwd  32 x 1961 tokens H=2 d=96", timeit(lambda: ops.attn_bwd(do3, qkv3, out3, lse3, lay10, 2)), 10.0 * D * lay10.sum_sq)
    # ---- row-wise
    gam, bet = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    report("layernorm fwd f32 -> bf16+f32", timeit(lambda: ops.layernorm_fwd(x32, gam, bet, 1e-5, out_f32=True)), 0, T * D * (4 + 2 + 4))
    report("layernorm2 fwd f32 -> f32 + bf16 (norm2 + next norm1)", timeit(lambda: ops.layernorm2_fwd(x32, gam, bet, 1e-5, gam, bet, 1e-5)), 0, T * D * (4 + 4 + 2))
    _, _, mean, rstd = ops.layernorm_fwd(x32, gam, bet, 1e-5)
    dg, db, dc = (torch.zeros(D, device=dev) for _ in range(3))
    report("layernorm bwd f32 (+dres) -> f32+bf16", timeit(lambda: ops.layernorm_bwd(x32, x32, gam, mean, rstd, dgamma=dg, dbeta=db, dcolsum=dc, dres=x32, want_bf16=True)), 0, T * D * (4 * 3 + 4 + 2))
    acc = torch.zeros(F, device=dev)
    report("colsum bf16 [T,2048]", timeit(lambda: ops.colsum(hid, acc)), 0, T * F * 2)
    acc = torch.zeros(3 * D, device=dev)
    report("colsum bf16 [T,576]", timeit(lambda: ops.colsum(qkv, acc)), 0, T * 3 * D * 2)
    n = 17_510_464
    p, gr, m, v, t = (torch.randn(n, device=dev) for _ in range(5))
    p16, t16 = torch.empty(n, device=dev, dtype=bf16), torch.empty(n, device=dev, dtype=bf16)
    report("adamw+ema+bf16 shadows 17.5M params", timeit(lambda: ops.adamw_step(p, gr, m, v, lr=1e-3, step=3, p_bf16=p16, teacher=t, teacher_bf16=t16, tau=0.99)), 0, n * 44.0)
    report("ema only 17.5M params (12 B/param)", timeit(lambda: ops.ema_update(t, p, 0.99)), 0, n * 12.0)


if __name__ == "__main__":
    main()
