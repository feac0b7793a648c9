#!/bin/bash
# Round 2, call 13 (one B200): where the half-form synthesis skeleton loses time (A-ring commits / hand-shakes removed one by one),
# 4 x 4-KS slots vs 2 x 8-KS slots; nle_mad on the device
mkdir -p gpurun_out
P=$PWD/cdlnet-video_b200
timeout -s KILL 200 python -m pytest tests/test_nle.py -q -x > gpurun_out/r02r_nle.log 2>&1; echo "nle rc=$?"; tail -4 gpurun_out/r02r_nle.log
for lib in libcdl_b200.so libcdl_b200_k8.so; do
  echo "== $lib phases (0 full, 448 skeleton, 960 skeleton without A ring, 1984 free-running MMAs)"
  CDL_LIB_PATH=$P/$lib timeout -s KILL 300 python scripts/syn_phase.py 4 0 64 448 960 1984 2>&1 | tail -1 | tee gpurun_out/r02r_phase_$lib.json
done
for lib in libcdl_b200_prof.so libcdl_b200_k8prof.so; do
  for mode in 0 448 960; do
    echo "== $lib timeline mode $mode"; CDL_TC_DBG_MODE=$mode CDL_LIB_PATH=$P/$lib timeout -s KILL 120 python scripts/tc_timeline.py 4 2>&1 | grep -A8 "== synthesis" | tee -a gpurun_out/r02r_timeline_$lib.log
  done
done
CDL_LIB_PATH=$P/libcdl_b200_k8.so timeout -s KILL 300 python -m pytest tests/test_tc_gpu.py -q -x 2>&1 | tail -3
CDL_LIB_PATH=$P/libcdl_b200_k8.so timeout -s KILL 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('k8 cfg2', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['per_kernel_ms_per_step'])"
