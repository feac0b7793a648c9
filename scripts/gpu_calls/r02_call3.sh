#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_tc_gpu.py -q -x > gpurun_out/r02c_tc.log 2>&1; echo "tc rc=$?"; tail -5 gpurun_out/r02c_tc.log
timeout -s KILL 200 python scripts/syn_phase.py 4 > gpurun_out/r02c_phase.json 2> gpurun_out/r02c_phase.err; echo "rc=$?"; cat gpurun_out/r02c_phase.json; tail -3 gpurun_out/r02c_phase.err
