#!/bin/bash
# Round-end style check on the B200 box: GPU parity tests, smoke, bench (both arms), ncu launch list.
# (the ncu --set full capture of the two tensor-core kernels is a separate call: scripts/gpu_ncu.sh)
# Usage: scripts/gpu_check.sh [tag]
TAG=${1:-r01}
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>/dev/null; echo "ref exit $?"; cut -c1-300 gpurun_out/${TAG}_bench_ref.json
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-breakdown"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches exit $?"
ls -la gpurun_out | tail -12
