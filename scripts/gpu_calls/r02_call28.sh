#!/bin/bash
# Round 2, call 28 (one B200): per-source-line profile of the 2-D synthesis kernel after the index-arithmetic cleanup
mkdir -p gpurun_out
export TC2_ARMS=tc2
ncu --set full --clock-control none --import-source on -k "regex:^k_tc2_synthesis" -s 20 -c 1 -f -o /tmp/ncu2d python scripts/tc2_bench.py cfg4 > gpurun_out/r02ag_ncu2d.log 2>&1; echo "ncu 2d exit $?"
python scripts/ncu_lines.py /tmp/ncu2d.ncu-rep k_tc2_synthesis 40 > gpurun_out/r02ag_lines_tc2_synthesis.txt 2>&1
python scripts/ncu_summary.py /tmp/ncu2d.ncu-rep > gpurun_out/r02ag_ncu2d_summary.txt 2>&1
cut -c1-200 gpurun_out/r02ag_lines_tc2_synthesis.txt | head -32; head -12 gpurun_out/r02ag_ncu2d_summary.txt | cut -c1-160; grep -E "issue_active|inst_executed.sum" gpurun_out/r02ag_ncu2d_summary.txt | cut -c1-160
