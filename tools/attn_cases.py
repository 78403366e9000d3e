#!/usr/bin/env python
"""Run attention-forward parity cases, each in its own process (a device trap poisons the CUDA context)."""
import subprocess
import sys

CASES = [  # (counts, npatch, D, H)
    ([7, 2, 2, 9, 1, 4], 196, 192, 12), ([7, 2, 2, 9, 1, 4], 196, 192, 2), ([10] * 40, 196, 192, 12), ([10] * 80, 196, 192, 2),
    ([1] * 400, 196, 192, 2), ([3] * 300, 196, 192, 2), ([1, 3] * 150, 196, 192, 2), ([7, 2, 2, 9, 1, 4] * 12, 196, 192, 2),
    ([7, 2, 2, 9, 1, 4] * 3, 196, 768, 12), ([7, 2, 2, 9, 1, 4] * 6, 196, 256, 2), ([5] * 200, 36, 192, 2),
]
CODE = r'''
import sys, math, torch
sys.path.insert(0, ".")
from chadavit_b200 import ops
from tests.test_attn_gpu import _ref_attn
counts, npatch, D, H = eval(sys.argv[1])
lay = ops.PackedLayout(counts, npatch, "cuda")
g = torch.Generator(device="cpu").manual_seed(1)
qkv = (torch.randn(lay.T, 3 * D, generator=g) * 1.5).to(torch.bfloat16).cuda()
out, lse = ops.attn_fwd(qkv, lay, H)
ops.sync_check()
ref, ref_lse = _ref_attn(qkv, lay.cu_host.tolist(), H)
print("err %.3e lse %.3e n_work %d" % ((out.float() - ref).abs().max().item(), (lse - ref_lse).abs().max().item(), lay.attn_work(H, 256).shape[0]))
'''
for c in CASES:
    try:
        r = subprocess.run([sys.executable, "-c", CODE, repr(c)], capture_output=True, text=True, timeout=120)
        tail = (r.stdout.strip().splitlines() or ["?"])[-1]
        print(f"{str(c)[:60]:60s} rc={r.returncode} {tail[:100]}")
    except subprocess.TimeoutExpired:
        print(f"{str(c)[:60]:60s} TIMEOUT")
