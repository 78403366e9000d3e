#!/bin/bash
# round 2: fused FFN backward (d(hidden) + dy) — parity, kernel A/B, step A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x -k "ffn_bwd" -s 2>&1 | grep -E "ffn bwd|passed|failed|Error|error" | head -20
timeout 300 python tools/microbench.py --only ffnbwd > gpurun_out/r2c_microbench_ffnbwd.txt 2>&1; cat gpurun_out/r2c_microbench_ffnbwd.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for n in off on off2 on2; do
  if [[ $n == off* ]]; then export CB_NO_FFN_BWD=1; else unset CB_NO_FFN_BWD; fi
  timeout 300 python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/r2c_bench_$n.json 2>/dev/null
done
python - <<'PY'
import json
for n in ("off", "on", "off2", "on2"):
    try:
        d = json.load(open(f"gpurun_out/r2c_bench_{n}.json"))
        a = d["roofline"]["all"]
        print(n, f"{d['value']:.1f} imgs/s  {d['ms_per_step']:.2f} ms  e2e {d['e2e']['value']:.1f}  gemm {a['cb_gemm_bf16']['ms_per_step']:.2f} ms  ffn_bwd {a.get('cb_ffn_bwd', {}).get('ms_per_step', 0):.2f} ms  clocks {d['clocks']['sm_mhz']}")
    except Exception as e:
        print(n, "failed", e)
PY
