#!/usr/bin/env python
"""CLS-query attention of the last block (cb_attn_cls_fwd / cb_attn_cls_bwd) timed alone on the two packed global crops of the bench
(128 sequences, T ~ 137 k tokens, 2 heads x 96): us and algorithmic GB/s (forward reads K, V: 4 T D bytes; backward + writes dqkv: 10 T D)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chadavit_b200 import ops  # noqa: E402
from tools.microbench import timeit  # noqa: E402

dev, bf16, D, H = "cuda", torch.bfloat16, 192, 2
counts = np.random.RandomState(1234).randint(1, 11, size=64).tolist()
for name, cts, npatch in (("global crops x2", counts + counts, 196), ("local crops x6", counts * 6, 36)):
    lay = ops.PackedLayout(cts, npatch, dev)
    qkv = (torch.randn(lay.T, 3 * D, device=dev) * 0.5).to(bf16)
    do = torch.randn(lay.B, D, device=dev).to(bf16)
    out, lse = ops.attn_cls_fwd(qkv, lay, H)
    tf = timeit(lambda: ops.attn_cls_fwd(qkv, lay, H))
    tb = timeit(lambda: ops.attn_cls_bwd(do, qkv, out, lse, lay, H))
    print(f"{name:16s} B={lay.B:4d} T={lay.T:7d}: fwd {tf:6.1f} us = {4.0 * lay.T * D / tf / 1e3:7.1f} GB/s   bwd {tb:6.1f} us = {10.0 * lay.T * D / tb / 1e3:7.1f} GB/s")
