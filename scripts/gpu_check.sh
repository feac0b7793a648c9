#!/bin/bash
# Run on the B200 box via gpurun: GPU parity tests, smoke, bench, then ncu launch list + one full capture.
# Usage: scripts/gpu_check.sh [tag] [ncu-kernel-regex]
TAG=${1:-r01}
KRE=${2:-k_cc_analysis}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/${TAG}_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
python bench.py --steps 1 --warmup 1 --clips 1 --no-cpu-baseline --no-breakdown > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --clips 1 --no-cpu-baseline --no-breakdown > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 2 -c 2 -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 1 --warmup 1 --clips 1 --no-cpu-baseline --no-breakdown > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out | tail -15
