mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x 2>&1 | tail -1
for r in 1 2; do for v in prev cur; do
  if [ $v == prev ]; then export CB_VARIANT=prev; else unset CB_VARIANT; fi
  echo "== $v run $r"; timeout 300 python tools/microbench.py 2>&1 | grep -E "^qkv|^proj|^du |^fc1|^fc2|^dy "
done; done
unset CB_VARIANT
for n in A1 B1 A2 B2; do
  if [[ $n == A* ]]; then export CB_VARIANT=prev; else unset CB_VARIANT; fi
  timeout 300 python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/step_ab_$n.json 2>/dev/null
done
python - <<'PY'
import json
for n in ("A1", "B1", "A2", "B2"):
    d = json.load(open(f"gpurun_out/step_ab_{n}.json")); a = d["roofline"]["all"]
    print(n, f"{d['value']:.1f} imgs/s  {d['ms_per_step']:.2f} ms  gemm {a['cb_gemm_bf16']['ms_per_step']:.2f}  clk {d['clocks']['sm_mhz']}")
PY
