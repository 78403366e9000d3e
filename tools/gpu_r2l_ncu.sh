#!/bin/bash
# ncu --set full of the out-projection + LayerNorm kernel and the fused FFN with the hidden store (the two largest HBM-side items of the step)
mkdir -p gpurun_out
for k in proj_ln ffn_save; do
  case $k in proj_ln) pat=gemm_kernel;; ffn_save) pat=ffn_kernel;; esac
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$pat" -s 2 -c 1 -f -o gpurun_out/r2l_ncu_$k python tools/prof_one.py $k > gpurun_out/r2l_ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
  python tools/ncu_pick.py gpurun_out/r2l_ncu_$k.ncu-rep > gpurun_out/r2l_ncu_$k.txt 2>&1
  cat gpurun_out/r2l_ncu_$k.txt
  ncu -i gpurun_out/r2l_ncu_$k.ncu-rep --page details --csv 2>/dev/null | grep -i -E "stall|Warp Cycles Per Issued|No Eligible|Theoretical Occupancy|L2 Hit|Mem Busy|Max Bandwidth" | cut -c1-260 | head -30 > gpurun_out/r2l_ncu_${k}_details.txt
  cat gpurun_out/r2l_ncu_${k}_details.txt
done
