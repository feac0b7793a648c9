#!/bin/bash
# Bring-up of the experimental 2-D tcgen05 analysis kernel (opt-in, CDL_TC2D=1) on the B200 box.
mkdir -p gpurun_out
export CDL_RUN_EXPERIMENTAL=1
timeout -s KILL ${1:-150} python -m pytest tests/test_tc2_gpu.py -q -s 2>&1 | tail -60 > gpurun_out/tc2_bringup.log
cat gpurun_out/tc2_bringup.log
timeout -s KILL ${2:-120} python scripts/tc2_bench.py ${3:-cfg1b cfg4 cfg3} 2>&1 | tail -20
