#!/bin/bash
# The 2-D tensor-core kernels on the B200 box: parity tests (incl. the opt-in variants), the 3-D smoke, per-config timing.
#   bash scripts/gpu_tc2.sh [test-timeout] [bench-timeout] [configs...]      TC2_ARMS=fp32,tc2[,tc2mp] selects the arms
mkdir -p gpurun_out
export CDL_RUN_EXPERIMENTAL=1
timeout -s KILL ${1:-150} python -m pytest tests/test_tc2_gpu.py -q -s 2>&1 | tail -60 > gpurun_out/tc2_bringup.log
cat gpurun_out/tc2_bringup.log
timeout -s KILL 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
T=${2:-120}; shift; shift
timeout -s KILL $T python scripts/tc2_bench.py ${@:-cfg1b cfg4 cfg3} 2>&1 | tail -20
