#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_backbone_gpu.py tests/test_attn_cls_gpu.py -q -s -k "fp32_parity or attn_cls" 2>&1 | grep -E "split3|attn_cls H=2 d=96|passed|failed|Error|assert" | tail -20
timeout 300 python tools/microbench_cls.py 2>&1 | tail -4
timeout 600 python bench.py --steps 10 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2i_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2i_bench.json"))
print(d["value"], d["ms_per_step"], d["parity_check"])
PY
