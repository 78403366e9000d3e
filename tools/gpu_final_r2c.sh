#!/bin/bash
# last state of round 2: full GPU suite + the complete bench line -> gpurun_out/final3_bench_1gpu.json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
[ "$1" == "--no-tests" ] || timeout 800 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/final3_bench_1gpu.json 2> gpurun_out/final3_bench_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/final3_bench_1gpu.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
for k in ("fixed_batch_graph", "multicrop_v8", "without_unused_local_crop_passes", "parity_check"):
    print(k, json.dumps(d.get(k))[:260])
PY
