#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_dino_gpu.py tests/test_next_rows_gpu.py tests/test_engine_r2_gpu.py tests/test_fullsize_gpu.py -q > gpurun_out/r2_t6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t6.log; tail -6 gpurun_out/r2_t6.log
for flag in "" "--serial-forward"; do
  timeout 600 python bench.py --no-extras --no-cpu-baseline $flag > gpurun_out/r2_bench_ab.log 2> gpurun_out/r2_bench_ab.err; echo "bench $flag rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_ab.log").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "fixed_graph", round(d["fixed_batch_graph"]["value"],1) if d.get("fixed_batch_graph") and "value" in d["fixed_batch_graph"] else d.get("fixed_batch_graph"), "tok", d["config"]["tokens_per_gpu_global_crop_mean"])
PY
done
