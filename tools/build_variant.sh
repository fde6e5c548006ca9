#!/bin/bash
# tools/build_variant.sh NAME -DFLAG...: the product library with extra defines on clip_loss_tc.cu ->
# tools/prof/libmae_clip_b200_NAME.so (loaded through MAE_CLIP_B200_LIB; timing experiments only)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
python -m mae_clip_b200._build > /dev/null
mkdir -p tools/prof
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -I include \
  -c mae_clip_b200/csrc/clip_loss_tc.cu -o tools/prof/clip_loss_tc_$name.o
objs=$(ls mae_clip_b200/build/*.o | grep -v clip_loss_tc.o)
nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o tools/prof/libmae_clip_b200_$name.so $objs tools/prof/clip_loss_tc_$name.o
echo tools/prof/libmae_clip_b200_$name.so
