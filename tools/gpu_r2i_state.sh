#!/bin/bash
# round 2: state of the step after the bias-gradient / fused FFN backward changes — tests, full bench line, launch list of one step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2i_bench_1gpu.json 2> gpurun_out/r2i_bench_1gpu.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 432 --csv --log-file gpurun_out/r2i_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2i_ncu_bench.log 2>&1
echo "ncu bench rc=$?"
python tools/ncu_summary.py gpurun_out/r2i_launches.csv > gpurun_out/r2i_launches.txt 2>&1; head -45 gpurun_out/r2i_launches.txt
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2i_bench_1gpu.json"))
print(f"value {d['value']:.1f} imgs/s  {d['ms_per_step']:.2f} ms  e2e {d['e2e']['value']:.1f}  fixed-graph {d['fixed_batch_graph']['value']:.1f}  launches {d['gpu_launches']}")
r = d["roofline"]; print("roofline", r["kernel"], round(r["achieved"], 1), round(r["frac"], 3), r["avg_launch_ms"])
for k, v in r["all"].items(): print("  ", k, round(v["ms_per_step"], 2), "ms", round(v["achieved_tflops"], 1), "TF", round(v["achieved_gbs"]), "GB/s")
print("parity", d.get("parity_check")); print("multicrop", d["multicrop_v8"]["value"], "cfg1", d["cfg1_extraction"]["imgs_per_s"], "cfg4", d["cfg4_attention_stress"]["fwd_tflops_per_gpu"], d["cfg4_attention_stress"]["bwd_tflops_per_gpu"])
PY
