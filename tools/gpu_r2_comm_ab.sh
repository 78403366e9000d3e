#!/bin/bash
# N GPUs (N = $1), one box: the gradient all-reduce variants back to back (bucketed 3 blocks = default | flat after the backward |
# 6-block buckets | 12-block bucket = one backbone all-reduce under nothing | NCCL limited to 4 CTAs), twice each.
cd "$(dirname "$0")/.."
N=$1
mkdir -p gpurun_out
run() {  # name, env, flags
  env $2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 --no-extras --no-cpu-baseline $3 > gpurun_out/comm_ab_$1.json 2>/dev/null
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/comm_ab_$1.json").read().strip().splitlines()[-1])
    print("$1", round(d["value"], 1), "imgs/s", round(d["ms_per_step"], 2), "ms  fixed-graph", round(d["fixed_batch_graph"]["ms_per_step"], 2), "ms  clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$1", "ERR", e)
PY
}
for r in 1 2; do
  run bucket3_$r "A=1" ""
  run flat_$r "A=1" "--no-overlap"
  run bucket6_$r "A=1" "--bucket-blocks 6"
  run maxctas4_$r "NCCL_MAX_CTAS=4" ""
done
