#!/bin/bash
# A/B of a head-GEMM environment switch: correctness with the switch at 1, then image-head timings at 0 / 1.
# Usage (GPU box): bash tools/head_gemm_ab.sh MAE_CLIP_GEMM_COALESCE > gpurun_out/head_gemm_ab.log 2>&1
set -u
VAR=${1:-MAE_CLIP_GEMM_COALESCE}
echo "== tests with $VAR=1"
env $VAR=1 timeout 100 python -m pytest tests/test_gpu_heads_model.py -x -q -m gpu 2>&1 | tail -3
for v in 0 1; do
  echo "== $VAR=$v"
  env $VAR=$v timeout 60 python tools/head_bench.py 32768 tc2048 2>&1 | grep "tc_f16x3"
done
