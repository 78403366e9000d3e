#!/usr/bin/env python
"""Launch ONE kernel class a few times so that `ncu --set full` can capture it in isolation.
    python tools/prof_one.py fc1|qkv|dh|fc2|attn_fwd|attn_bwd|ffn|ffn_save|ffn_bwd|ln_bwd|proj_ln|attn_cls_fwd|attn_cls_bwd"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chadavit_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "fc1"
T, D, F, dev, bf16 = 68664, 192, 2048, "cuda", torch.bfloat16
r = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(bf16)  # noqa: E731
x16, hid, x32 = r(T, D), r(T, F), torch.randn(T, D, device=dev)
w_in, w1, w2 = r(3 * D, D), r(F, D), r(D, F)
b1, b2, b_in = torch.randn(F, device=dev), torch.randn(D, device=dev), torch.randn(3 * D, device=dev)
counts = np.random.RandomState(1234).randint(1, 11, size=64).tolist()
lay = ops.PackedLayout(counts, 196, dev)
qkv = r(lay.T, 3 * D)
do = r(lay.T, D)
R = ops.EPI_RESIDUAL_F32 | ops.EPI_OUT_F32
fns = {
    "fc1": lambda: ops.gemm(x16, w1, bias=b1, flags=ops.EPI_RELU),
    "qkv": lambda: ops.gemm(x16, w_in, bias=b_in),
    "dh": lambda: ops.gemm(x16, w2, b_mn=True, aux=hid, flags=ops.EPI_RELU_MASK),
    "fc2": lambda: ops.gemm(hid, w2, bias=b2, aux=x32, flags=R),
    "attn_fwd": lambda: ops.attn_fwd(qkv, lay, 2),
    "ffn": lambda: ops.ffn_fwd(x16, w1, b1, w2, b2, x32, save_hidden=False),
    "ffn_save": lambda: ops.ffn_fwd(x16, w1, b1, w2, b2, x32, save_hidden=True),
}
if which == "ffn_bwd":
    bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (F // 32, (T + 31) // 32 * 32), device=dev, dtype=torch.int32)
    fns["ffn_bwd"] = lambda: ops.ffn_bwd(x16, w2, w1, bits, x32)
if which == "ln_bwd":
    gam = torch.ones(D, device=dev)
    _, _, mean, rstd = ops.layernorm_fwd(x32, gam, torch.zeros(D, device=dev), 1e-5)
    dy32, dres32 = torch.randn(T, D, device=dev), torch.randn(T, D, device=dev)
    acc = [torch.zeros(D, device=dev) for _ in range(3)]
    fns["ln_bwd"] = lambda: ops.layernorm_bwd(dy32, x32, gam, mean, rstd, dgamma=acc[0], dbeta=acc[1], dcolsum=acc[2], dres=dres32, want_bf16=True)
if which == "attn_bwd":
    out, lse = ops.attn_fwd(qkv, lay, 2)
    fns["attn_bwd"] = lambda: ops.attn_bwd(do, qkv, out, lse, lay, 2)
if which == "proj_ln":
    w_o, b_o, gam, bet = r(D, D), torch.randn(D, device=dev), torch.ones(D, device=dev), torch.zeros(D, device=dev)
    fns["proj_ln"] = lambda: ops.gemm_ln_fwd(x16, w_o, b_o, x32, gam, bet, 1e-5, keep_z=True)
if which in ("attn_cls_fwd", "attn_cls_bwd"):
    lay2 = ops.PackedLayout(counts + counts, 196, dev)        # the two packed global crops, as the engine runs them
    qkv2, do2 = r(lay2.T, 3 * D), r(lay2.B, D)
    oc, lc = ops.attn_cls_fwd(qkv2, lay2, 2)
    fns["attn_cls_fwd"] = lambda: ops.attn_cls_fwd(qkv2, lay2, 2)
    fns["attn_cls_bwd"] = lambda: ops.attn_cls_bwd(do2, qkv2, oc, lc, lay2, 2)
for _ in range(5):
    fns[which]()
torch.cuda.synchronize()
