#!/bin/bash
# Round 2, call 43 (one B200): ncu --set full of the final 2-D kernels (config 4), summarised on the box
mkdir -p gpurun_out
export TC2_ARMS=tc2
ncu --set full --clock-control none -k "regex:^k_tc2_(analysis|synthesis)" -s 40 -c 2 -f -o /tmp/ncu2d python scripts/tc2_bench.py cfg4 > gpurun_out/r02av_ncu2d.log 2>&1; echo "ncu exit $?"
python scripts/ncu_summary.py /tmp/ncu2d.ncu-rep > gpurun_out/r02av_ncu2d_summary.txt 2>&1; head -8 gpurun_out/r02av_ncu2d_summary.txt | cut -c1-170; grep -E "issue_active|inst_executed.sum" gpurun_out/r02av_ncu2d_summary.txt | cut -c1-170
