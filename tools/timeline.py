#!/usr/bin/env python
"""In-kernel timeline of the attention kernels (debug build only).

    CB_NVCC_EXTRA=-DCB_TIMELINE python -m chadavit_b200.build --force
    python tools/timeline.py fwd|bwd [uniform|ragged] > gpurun_out/timeline_fwd.txt

CTA 0 records (clock64, tag) per role (0 = MMA issuer, 1/2 = lane 0 of the first softmax warp of each warpgroup); this
script prints the events of the first work item merged in time order, plus per-phase average durations."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chadavit_b200 import _lib, ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
shape = sys.argv[2] if len(sys.argv) > 2 else "uniform"
dev, bf16, D = "cuda", torch.bfloat16, 192
counts = [10] * 32 if shape == "uniform" else np.random.RandomState(1234).randint(1, 11, size=64).tolist()
lay = ops.PackedLayout(counts, 196, dev)
r = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(bf16)  # noqa: E731
qkv, do = r(lay.T, 3 * D), r(lay.T, D)
lib = _lib.load()
ROLES, LEN = 4, 4096
buf = np.zeros((ROLES, LEN), dtype=np.uint64)
rd = getattr(lib, "cb_debug_timeline_" + {"ffnstore": "ffn", "ffnbwd": "ffn"}.get(which, which))
rd.argtypes, rd.restype = [C.c_void_p], C.c_int

out, lse = ops.attn_fwd(qkv, lay, 2)
T, F = 68664, 2048
x16, w1, b1 = r(T, D), r(F, D), torch.randn(F, device=dev)
hid_out = torch.empty(T, F, device=dev, dtype=bf16)
w2f, b2f, x32f = r(D, F), torch.randn(D, device=dev), torch.randn(T, D, device=dev)
tl_bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (F // 32, (T + 31) // 32 * 32), device=dev, dtype=torch.int32)


def run():
    if which == "fwd":
        ops.attn_fwd(qkv, lay, 2)
    elif which in ("bwd", "bwd2"):
        ops.attn_bwd(do, qkv, out, lse, lay, 2)
    elif which in ("ffn", "ffn2", "ffn3"):
        ops.ffn_fwd(x16, w1, b1, w2f, b2f, x32f, save_hidden=False)
    elif which == "ffnstore":       # pair-of-tiles kernel with the hidden store + mask bits (student pass)
        ops.ffn_fwd(x16, w1, b1, w2f, b2f, x32f, save_hidden=True, save_mask_bits=True, kernel=1)
    elif which == "ffnbwd":         # fused backward through the hidden layer (same translation unit, same tags)
        ops.ffn_bwd(x16, w2f, w1, tl_bits, x32f)
    else:   # gemm: fc1
        ops.gemm(x16, w1, bias=b1, flags=ops.EPI_RELU, out=hid_out)


for _ in range(3):
    run()
torch.cuda.synchronize()
rd(buf.ctypes.data)  # discard warm-up
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
run()
e1.record()
torch.cuda.synchronize()
print(f"# {which} {shape}: {e0.elapsed_time(e1) * 1e3:.1f} us (instrumented)")
rd(buf.ctypes.data)
ev = []
for role in range(ROLES):
    for v in buf[role]:
        if v == 0:
            break
        ev.append((int(v >> np.uint64(8)), role, int(v & np.uint64(255))))
ev.sort()
t0 = ev[0][0]
# optional window: python tools/timeline.py <which> <shape> <first_cycle> <last_cycle>  prints every event in that range instead
win = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else None
print("# cycles role tag   " + (f"(window {win[0]}..{win[1]})" if win else "(first 160 events)"))
for t, role, tag in (ev[:160] if win is None else [e for e in ev if win[0] <= e[0] - t0 <= win[1]]):
    print(f"{t - t0:9d}  {'   ' * role}r{role}:{tag}")
# per role: average delta between consecutive events keyed by (tag_prev -> tag)
for role in range(ROLES):
    seq = [(t, tag) for t, rr, tag in ev if rr == role]
    if len(seq) < 2:
        continue
    agg = {}
    for (ta, ga), (tb, gb) in zip(seq[:-1], seq[1:]):
        k = (ga, gb)
        a = agg.setdefault(k, [0, 0])
        a[0] += 1
        a[1] += tb - ta
    print(f"# role {role}: {len(seq)} events, span {seq[-1][0] - seq[0][0]} cycles")
    for k, (n, s) in sorted(agg.items()):
        print(f"#   {k[0]:3d} -> {k[1]:3d}: n={n:5d} avg={s / n:9.1f} total={s}")
