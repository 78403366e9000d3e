#!/bin/bash
# 2 GPUs: NCCL equivalence test + 2-rank bench (overlapped vs flat all-reduce)
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_engine_r2_gpu.py -x -q > gpurun_out/r2_t4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t4.log; tail -8 gpurun_out/r2_t4.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-extras > gpurun_out/r2_bench_2gpu.log 2> gpurun_out/r2_bench_2gpu.err; echo "bench2 rc=$?"; tail -c 600 gpurun_out/r2_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-extras --no-overlap > gpurun_out/r2_bench_2gpu_flat.log 2> gpurun_out/r2_bench_2gpu_flat.err; echo "bench2 flat rc=$?"
python - <<PY
import json
for f in ("gpurun_out/r2_bench_2gpu.log","gpurun_out/r2_bench_2gpu_flat.log"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "fixed", d["fixed_batch_graph"], d["config"]["rank_work_spread"], d["config"]["grad_allreduce"])
    except Exception as e:
        print(f, "ERR", e)
PY
