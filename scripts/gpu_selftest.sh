#!/bin/bash
# tcgen05 primitive bring-up on the B200 box
mkdir -p gpurun_out
T=cdlnet-video_b200/csrc/selftest/tc_selftest
{
for cfg in "1 0 176 7" "1 1 176 7" "1 1 64 2" "2 0 176 7" "2 0 192 7" "2 1 192 7" "2 1 176 7" "2 1 256 4" "1 0 256 43"; do
  echo "== $cfg"; timeout 30 $T $cfg; echo "rc=$?"
done
} > gpurun_out/selftest.log 2>&1
cat gpurun_out/selftest.log
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -15
python __graft_entry__.py smoke 2>&1 | tail -3
# issue-cost variants (built on the CPU box):
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I cdlnet-video_b200/csrc [-DSELFTEST_WARP_ISSUE [-DSELFTEST_CONST]] \
#        -o cdlnet-video_b200/csrc/selftest/tc_selftest[_warp|_const] cdlnet-video_b200/csrc/selftest/tc_selftest.cu -lcuda
#   args: cta_group ts N ksteps rep probe commit_every overlap   (overlap = 1: SS form through the LBO 16 / SBO 144 overlapping descriptor)
