#!/bin/bash
# round-2 evidence: ncu --set full of the attention backward (generation 2) and forward, launch list of one bench step
cd "$(dirname "$0")/.."
for k in attn_bwd attn_fwd; do
  pat=$([ $k == attn_bwd ] && echo "attn_bwd2_kernel" || echo "attn_fwd2_kernel")
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 -f -o gpurun_out/prof_r02_$k python tools/prof_one.py $k > gpurun_out/prof_r02_$k.log 2>&1
  echo "ncu $k rc=$?"
done
# launch list of one eager bench step (fresh batches, no extras): kernel shares
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4200 -c 460 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench_r2.log 2>&1
echo "ncu bench rc=$?"
python tools/ncu_summary.py gpurun_out/launches_r2.csv > gpurun_out/launches_r2.txt 2>&1; head -30 gpurun_out/launches_r2.txt
