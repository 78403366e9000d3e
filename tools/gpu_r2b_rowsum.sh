#!/bin/bash
# round 2: bias gradients on the tensor pipe of the weight-gradient products — parity, kernel A/B, step A/B
mkdir -p gpurun_out
python -m pytest tests/test_gemm_gpu.py -q -x -k "weight_grad or epilogues or mask" 2>&1 | tail -5
python tools/microbench.py --only dw > gpurun_out/r2b_microbench_dw.txt 2>&1; cat gpurun_out/r2b_microbench_dw.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
CB_NO_GEMM_ROWSUM=1 python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/r2b_bench_off.json 2>/dev/null
python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/r2b_bench_on.json 2>/dev/null
CB_NO_GEMM_ROWSUM=1 python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/r2b_bench_off2.json 2>/dev/null
python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/r2b_bench_on2.json 2>/dev/null
python - <<'PY'
import json
for n in ("off", "on", "off2", "on2"):
    d = json.load(open(f"gpurun_out/r2b_bench_{n}.json"))
    a = d["roofline"]["all"]
    print(n, f"{d['value']:.1f} imgs/s  {d['ms_per_step']:.2f} ms  e2e {d['e2e']['value']:.1f}  gemm {a['cb_gemm_bf16']['ms_per_step']:.2f} ms  top {d['roofline']['kernel'][:24]} frac {d['roofline']['frac']:.3f}  clocks {d['clocks']['sm_mhz']}")
PY
