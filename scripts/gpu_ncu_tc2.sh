#!/bin/bash
# ncu full capture of the 2-D tensor-core kernels at config 4 (GDLNet 64 x 3 x 512^2), after a plain run of the same
# command; then the launch list of a whole forward.  Read with scripts/ncu_summary.py gpurun_out/<tag>_prof.ncu-rep.
#   bash scripts/gpu_ncu_tc2.sh [tag] [config] [arms]        e.g.  bash scripts/gpu_ncu_tc2.sh ncu2d cfg4 tc2v2
TAG=${1:-ncu2d}; CFG=${2:-cfg4}; export TC2_ARMS=${3:-tc2}
mkdir -p gpurun_out
CMD="python scripts/tc2_bench.py $CFG"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:^k_tc2_(analysis|synthesis)" -s 40 -c 2 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/${TAG}_ncu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list exit $?"
