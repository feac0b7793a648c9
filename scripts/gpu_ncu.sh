#!/bin/bash
# ncu full capture of the two tensor-core kernels (bench workload: 4 clips), after a plain run of the same command
TAG=${1:-ncu}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --clips 4 --no-cpu-baseline --no-breakdown"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:^k_tc_(analysis|synthesis)" -s 5 -c 4 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/${TAG}_ncu.log
