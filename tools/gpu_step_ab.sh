#!/bin/bash
# Same-box ABAB of the whole step with one environment switch: bash tools/gpu_step_ab.sh VAR VALUE  (A = VAR=VALUE, B = unset)
mkdir -p gpurun_out
VAR=$1; VAL=$2
for n in A1 B1 A2 B2; do
  if [[ $n == A* ]]; then export $VAR=$VAL; else unset $VAR; fi
  timeout 300 python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/step_ab_$n.json 2>/dev/null
done
unset $VAR
python - <<'PY'
import json
for n in ("A1", "B1", "A2", "B2"):
    try:
        d = json.load(open(f"gpurun_out/step_ab_{n}.json")); a = d["roofline"]["all"]
        print(n, f"{d['value']:.1f} imgs/s  {d['ms_per_step']:.2f} ms  e2e {d['e2e']['value']:.1f}  " + "  ".join(f"{k.replace('cb_', '')} {v['ms_per_step']:.2f}" for k, v in a.items()) + f"  clk {d['clocks']['sm_mhz']}")
    except Exception as e:
        print(n, "failed", e)
PY
