#!/bin/bash
mkdir -p gpurun_out
CDL_LIB_PATH=$PWD/cdlnet-video_b200/libcdl_b200_prof.so timeout -s KILL 120 python scripts/tc_timeline.py 4 > gpurun_out/r02h_timeline.log 2>&1; echo "rc=$?"; cat gpurun_out/r02h_timeline.log
timeout -s KILL 200 python scripts/syn_phase.py 16 0 448 2>&1 | tail -2
timeout -s KILL 200 python scripts/syn_phase.py 1 0 448 2>&1 | tail -2
