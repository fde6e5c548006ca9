#!/bin/bash
# A/B of the A-resident short-K GEMM form (MAE_CLIP_GEMM_ARES): correctness first, then head timings.
# Usage (GPU box): bash tools/head_ares_ab.sh > gpurun_out/head_ares_ab.log 2>&1
set -u
echo "== tests with MAE_CLIP_GEMM_ARES=1"
MAE_CLIP_GEMM_ARES=1 timeout 100 python -m pytest tests/test_gpu_heads_model.py -x -q -m gpu -k "tc_gemm or head_vs_oracle" 2>&1 | tail -3
for cfg in "0 0" "1 0" "1 4"; do
  set -- $cfg
  echo "== ARES=$1 GROUP=$2"
  if [ "$2" = "0" ]; then
    MAE_CLIP_GEMM_ARES=$1 timeout 60 python tools/head_bench.py 32768 tc2048 2>&1 | grep "E=2048 tc_f16x3"
  else
    MAE_CLIP_GEMM_ARES=$1 MAE_CLIP_GEMM_ARES_GROUP=$2 timeout 60 python tools/head_bench.py 32768 tc2048 2>&1 | grep "E=2048 tc_f16x3"
  fi
done
