#!/usr/bin/env python
"""Where does the end-to-end gap come from?  One process, one box, the bench's model and batches:
resident inputs without / with a blocking loss read per step, staged host inputs with a blocking read (bench.py's e2e) and with the
read lagging one step (the loss of step i is read while step i+1 runs; every loss is still read inside the timed region)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from chadavit_b200.methods import DINO  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
torch.manual_seed(0)
model = DINO(bench.dino_cfg(False, graph=False)).to(dev)
pools, host = bench.make_pools(1234, dev), bench.make_pools(1234, dev, pin=True)
K = 20
counts = [bench.step_counts(t, 0, 1) for t in range(4 * K + 40)]
cur = [0]


def nb(src):
    cur[0] += 1
    c = counts[cur[0] % len(counts)]
    return ([p[:sum(c)] for p in src], None, [c] * 8)


model.fused_train_step(([p for p in pools], None, [[10] * 64] * 8))
for _ in range(8):
    model.fused_train_step(nb(pools))
for _ in range(3):
    model.fused_train_step(model.stage_batch(nb(host))).item()


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


def resident(read):
    for _ in range(K):
        l = model.fused_train_step(nb(pools))
        if read:
            l.item()


def staged(lag):
    nxt, prev = model.stage_batch(nb(host)), None
    for i in range(K):
        l = model.fused_train_step(nxt)
        if i + 1 < K:
            nxt = model.stage_batch(nb(host))
        if lag:
            if prev is not None:
                prev.item()
            prev = l
        else:
            l.item()
    if lag:
        prev.item()


def resident_lag():
    prev = None
    for _ in range(K):
        l = model.fused_train_step(nb(pools))
        if prev is not None:
            prev.item()
        prev = l
    prev.item()


def staged_noread():
    nxt = model.stage_batch(nb(host))
    for i in range(K):
        model.fused_train_step(nxt)
        if i + 1 < K:
            nxt = model.stage_batch(nb(host))


import time  # noqa: E402
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(K):
    model.fused_train_step(nb(pools))
t_enq = (time.perf_counter() - t0) / K * 1e3
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(K):
    model.stage_batch(nb(host))
t_stage = (time.perf_counter() - t0) / K * 1e3
torch.cuda.synchronize()
print(f"host time to enqueue one step {t_enq:.2f} ms (GPU not waited for); host time of stage_batch {t_stage:.2f} ms")
def staged_async():
    nxt = model.stage_batch(nb(host))
    for i in range(K):
        l = model.fused_train_step(nxt, loss_to_host=True)
        if i + 1 < K:
            nxt = model.stage_batch(nb(host))
        l.item()


def resident_async():
    for _ in range(K):
        model.fused_train_step(nb(pools), loss_to_host=True).item()


for rep in range(2):
    print(f"rep {rep}: resident + loss_to_host read {timed(resident_async):.2f} ms | staged + loss_to_host read {timed(staged_async):.2f} ms")
for rep in range(2):
    print(f"rep {rep}: resident + lagged read {timed(resident_lag):.2f} ms | staged, no read {timed(staged_noread):.2f} ms")
for rep in range(2):
    print(f"rep {rep}: resident, no read {timed(lambda: resident(False)):.2f} ms | resident + blocking read {timed(lambda: resident(True)):.2f} ms | "
          f"staged + blocking read {timed(lambda: staged(False)):.2f} ms | staged + lagged read {timed(lambda: staged(True)):.2f} ms")
