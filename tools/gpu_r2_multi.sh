#!/bin/bash
# N GPUs (N = $1): NCCL equivalence tests (N = 2 only) + the N-rank bench line -> gpurun_out/final2_bench_${N}gpu.json
cd "$(dirname "$0")/.."
N=$1
mkdir -p gpurun_out
if [ "$N" == 2 ]; then
  timeout 600 python -m pytest tests/test_engine_r2_gpu.py tests/test_attn_cls_gpu.py -q 2>&1 | tail -2
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/final2_bench_${N}gpu.json 2> gpurun_out/final2_bench_${N}gpu.err; echo "bench$N rc=$?"; tail -c 400 gpurun_out/final2_bench_${N}gpu.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/final2_bench_${N}gpu.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "fixed", d["fixed_batch_graph"], d["config"]["rank_work_spread"], d.get("multicrop_v8"), d.get("cfg4_attention_stress"))
except Exception as e:
    print("ERR", e)
PY
