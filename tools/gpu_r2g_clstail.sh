#!/bin/bash
# CLS-only tail of the last block: kernel parity, model parity, then a same-box ABAB of the whole step (A = dense last block).
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_attn_cls_gpu.py tests/test_backbone_gpu.py -q -s 2>&1 | grep -E "attn_cls H=2 d=96|CLS-only|passed|failed|Error|error|assert" | tail -40
  timeout 900 python -m pytest tests/test_dino_gpu.py tests/test_next_rows_gpu.py tests/test_fullsize_gpu.py tests/test_engine_r2_gpu.py -q 2>&1 | tail -3 ) > gpurun_out/r2g_clstail_tests.txt 2>&1
cat gpurun_out/r2g_clstail_tests.txt
bash tools/gpu_step_ab.sh CB_NO_CLS_TAIL 1 > gpurun_out/r2g_clstail_ab.txt 2>&1
cat gpurun_out/r2g_clstail_ab.txt
