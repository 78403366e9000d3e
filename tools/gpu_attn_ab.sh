#!/bin/bash
# attention backward generation 2: parity tests, A/B against generation 1, in-kernel timeline
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_attn_gpu.py -x -q > gpurun_out/r2_t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t2.log
tail -3 gpurun_out/r2_t2.log
echo "--- generation 2"; timeout 300 python tools/microbench.py --only attn 2>&1 | tee gpurun_out/r2_attn_v2.txt
if [ "$1" == "ab" ]; then echo "--- generation 1"; CB_ATTN_BWD_V=1 timeout 300 python tools/microbench.py --only attn 2>&1 | tee gpurun_out/r2_attn_v1.txt; fi
CB_VARIANT=tl timeout 300 python tools/timeline.py bwd2 uniform > gpurun_out/r2_tl_bwd2_uniform.txt 2>&1; grep "^#" gpurun_out/r2_tl_bwd2_uniform.txt
