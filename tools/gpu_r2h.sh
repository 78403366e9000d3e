#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; tail -25 gpurun_out/r2h_bench.err
timeout 600 python -m pytest tests/test_backbone_gpu.py -q -s -k "tail" 2>&1 | grep -E "tail|CLS-only|passed|failed|Error|assert" | tail -30
