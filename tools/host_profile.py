#!/usr/bin/env python
"""cProfile of the HOST side of fused_train_step on fresh ragged batches (GPU not waited for inside the profiled region)."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from chadavit_b200.methods import DINO  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
torch.manual_seed(0)
model = DINO(bench.dino_cfg(False, graph=False)).to(dev)
pools = bench.make_pools(1234, dev)
counts = [bench.step_counts(t, 0, 1) for t in range(40)]
batches = [([p[:sum(c)] for p in pools], None, [c] * 8) for c in counts]
model.fused_train_step(([p for p in pools], None, [[10] * 64] * 8))
for b in batches[:8]:
    model.fused_train_step(b)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for b in batches[8:18]:
    model.fused_train_step(b)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(38)
st.sort_stats("cumulative").print_stats(30)
