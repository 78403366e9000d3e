#!/bin/bash
# round 2, final state: smoke(), GPU tests, full bench line, launch list of one step, ncu --set full of the attention kernels and the
# fused FFN backward (summaries -> gpurun_out/final_*; copied to profiles/ by hand)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_bench_reference_arm.json 2>/dev/null; echo "reference arm rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3650 -c 396 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/final_ncu_bench.log 2>&1
echo "ncu launch list rc=$?"
python tools/ncu_summary.py gpurun_out/final_launches.csv > gpurun_out/final_launches.txt 2>&1
for k in attn_bwd attn_fwd ffn_bwd ln_bwd; do
  case $k in attn_bwd) pat=attn_bwd2_kernel;; attn_fwd) pat=attn_fwd2_kernel;; ffn_bwd) pat=ffn_kernel;; ln_bwd) pat=layernorm_bwd_g16_kernel;; esac
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$pat" -s 2 -c 1 -f -o gpurun_out/final_ncu_$k python tools/prof_one.py $k > gpurun_out/final_ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
  python tools/ncu_pick.py gpurun_out/final_ncu_$k.ncu-rep > gpurun_out/final_ncu_$k.txt 2>&1
done
head -30 gpurun_out/final_launches.txt
cat gpurun_out/final_ncu_attn_bwd.txt
