#!/bin/bash
# Round 2, call 35 (one B200): does the tcgen05 tf32 MMA rate depend on how many distinct operand pairs the loop streams from shared memory?
S=cdlnet-video_b200/csrc/selftest
mkdir -p gpurun_out
{
for ks in 7 22 30; do
  for nc in 1 74; do
    echo "== SS (overlapping A descriptor), cta_group 2, $ks distinct k-steps, clusters=$nc"; timeout 60 $S/tc_selftest_const_ks$ks 2 0 176 $ks 600 0 0 1 $nc 2>&1 | grep -E "^timing|error"
  done
done
echo "== TS, 22 k-steps, 74 clusters"; timeout 60 $S/tc_selftest_const_ks22 2 1 176 22 600 0 0 1 74 2>&1 | grep -E "^timing|error"
echo "== cta_group 1 SS, 22 k-steps, 148 CTAs"; timeout 60 $S/tc_selftest_const_ks22 1 0 176 22 600 0 0 1 148 2>&1 | grep -E "^timing|error"
} | tee gpurun_out/r02an_mma_rate_vs_operand_footprint.log
