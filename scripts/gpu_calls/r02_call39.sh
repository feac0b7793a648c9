#!/bin/bash
# Round 2, call 39 (one B200): analysis kernel store L2 policy (evict-first default / evict-normal / evict-last) and L2 prefetch distance 1
mkdir -p gpurun_out
P=$PWD/cdlnet-video_b200
for arm in base st1 st2 pf1 base st1 st2 pf1; do
  lib=$P/libcdl_b200.so; [ $arm != base ] && lib=$P/libcdl_b200_$arm.so
  echo "== $arm"; CDL_LIB_PATH=$lib timeout -s KILL 200 python scripts/syn_phase.py 16 0 2>&1 | tail -1
done | tee gpurun_out/r02ar_ana_policy_ab.txt
