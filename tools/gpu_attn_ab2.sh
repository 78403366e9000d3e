#!/bin/bash
# Same-box A/B of two builds of the library: lib (current tree) vs lib_prev, e.g. built from the previous commit with
#   git stash; CB_VARIANT=prev python -m chadavit_b200.build; git stash pop; python -m chadavit_b200.build
# Box-to-box variance of a microbenchmark is ~3 %: only alternating runs on ONE box separate a 1 % change from noise
# (round 2: the split K / V barriers + early dV epilogue of attn_bwd2 looked better in the in-kernel timeline and measured
# 1.2-1.5 % slower here: 267.3 vs 264.2 us ragged, 519 vs 511.7 us for the two packed crops; not kept).
mkdir -p gpurun_out
SECTION=${1:-attn}
for r in 1 2 3; do
  for v in prev cur; do
    if [ $v == prev ]; then export CB_VARIANT=prev; else unset CB_VARIANT; fi
    echo "== $v run $r"; timeout 200 python tools/microbench.py --only $SECTION 2>&1
  done
done
