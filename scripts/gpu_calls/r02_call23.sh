#!/bin/bash
# Round 2, call 23 (one B200): per-source-line stall / instruction profiles of the four tensor-core kernels (summarised on the box)
mkdir -p gpurun_out
export TC2_ARMS=tc2
ncu --set full --clock-control none --import-source on -k "regex:^k_tc2_(analysis|synthesis)" -s 40 -c 2 -f -o /tmp/ncu2d python scripts/tc2_bench.py cfg4 > gpurun_out/r02ab_ncu2d.log 2>&1; echo "ncu 2d exit $?"
python scripts/ncu_lines.py /tmp/ncu2d.ncu-rep k_tc2_synthesis 45 > gpurun_out/r02ab_lines_tc2_synthesis.txt 2>&1
python scripts/ncu_lines.py /tmp/ncu2d.ncu-rep k_tc2_analysis 30 > gpurun_out/r02ab_lines_tc2_analysis.txt 2>&1
CMD2="python bench.py --workload cfg2 --steps 1 --warmup 1 --clips 4 --no-cpu-baseline --no-breakdown --no-e2e"
ncu --set full --clock-control none --import-source on -k "regex:^k_tc_(analysis|synthesis)" -s 10 -c 2 -f -o /tmp/ncu3d $CMD2 > gpurun_out/r02ab_ncu3d.log 2>&1; echo "ncu 3d exit $?"
python scripts/ncu_lines.py /tmp/ncu3d.ncu-rep k_tc_synthesis 45 > gpurun_out/r02ab_lines_tc_synthesis.txt 2>&1
python scripts/ncu_lines.py /tmp/ncu3d.ncu-rep k_tc_analysis 40 > gpurun_out/r02ab_lines_tc_analysis.txt 2>&1
head -30 gpurun_out/r02ab_lines_tc2_synthesis.txt | cut -c1-230
