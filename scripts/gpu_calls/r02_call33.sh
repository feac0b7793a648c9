#!/bin/bash
# Round 2, call 33 (one B200): tcgen05 tf32 MMA rate (M = 128 x cta_group, N = 176, K = 8, constant operands, converged-warp issue) on one
# cluster vs on the whole chip at once - is the ~118 cycles per MMA seen inside the kernels a chip-level limit?
T=cdlnet-video_b200/csrc/selftest/tc_selftest_const
mkdir -p gpurun_out
{
for nc in 1 8 37 74; do
  for ts in 1 0; do
    echo "== cta_group 2, ts=$ts (0 = SS through the overlapping descriptor), clusters=$nc"; timeout 60 $T 2 $ts 176 7 2000 0 0 1 $nc 2>&1 | grep -E "timing"
  done
done
for nc in 1 148; do echo "== cta_group 1, SS, CTAs=$nc"; timeout 60 $T 1 0 176 7 2000 0 0 1 $nc 2>&1 | grep -E "timing"; done
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader
} | tee gpurun_out/r02al_mma_rate_whole_chip.log
