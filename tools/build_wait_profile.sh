#!/bin/bash
# Builds tools/prof/libmae_clip_b200_prof.so: the product library with -DMC_WAIT_PROFILE on clip_loss_tc.cu (cycle counters
# around every class of mbarrier wait in pair_kernel).  Used by tools/wait_profile.py through MAE_CLIP_B200_LIB.
set -e
cd "$(dirname "$0")/.."
python -m mae_clip_b200._build > /dev/null
mkdir -p tools/prof
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DMC_WAIT_PROFILE -I include \
  -c mae_clip_b200/csrc/clip_loss_tc.cu -o tools/prof/clip_loss_tc.o
objs=$(ls mae_clip_b200/build/*.o | grep -v clip_loss_tc.o)
nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o tools/prof/libmae_clip_b200_prof.so $objs tools/prof/clip_loss_tc.o
echo tools/prof/libmae_clip_b200_prof.so
