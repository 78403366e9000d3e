#!/bin/bash
# round 2, final state (after the CLS-only last block): smoke(), GPU tests, full bench line, reference arm, launch list of one step,
# ncu --set full of the attention backward (the roofline kernel) and the two CLS attention kernels -> gpurun_out/final2_*
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/final2_bench_1gpu.json 2> gpurun_out/final2_bench_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/final2_bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final2_bench_reference_arm.json 2>/dev/null; echo "reference arm rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/final2_launches_all.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/final2_ncu_bench.log 2>&1
echo "ncu launch list rc=$?"
python tools/ncu_summary.py gpurun_out/final2_launches_all.csv --one-step > gpurun_out/final2_launches.txt 2>&1; head -12 gpurun_out/final2_launches.txt
for k in attn_bwd attn_cls_fwd attn_cls_bwd qkv; do
  case $k in attn_bwd) pat=attn_bwd2_kernel;; attn_cls_fwd) pat=attn_cls_fwd_kernel;; attn_cls_bwd) pat=attn_cls_bwd_kernel;; qkv) pat=gemm_kernel;; esac
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$pat" -s 2 -c 1 -f -o gpurun_out/final2_ncu_$k python tools/prof_one.py $k > gpurun_out/final2_ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
  python tools/ncu_pick.py gpurun_out/final2_ncu_$k.ncu-rep > gpurun_out/final2_ncu_$k.txt 2>&1
done
python - <<'PY'
import json
d = json.load(open("gpurun_out/final2_bench_1gpu.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
for k in ("fixed_batch_graph", "multicrop_v8", "without_unused_local_crop_passes", "cfg1_extraction", "parity_check"):
    print(k, json.dumps(d.get(k))[:400])
PY
cat gpurun_out/final2_ncu_attn_cls_fwd.txt gpurun_out/final2_ncu_attn_cls_bwd.txt
