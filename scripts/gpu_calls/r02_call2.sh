#!/bin/bash
# Round 2, GPU call 2: the redesigned video synthesis kernel (TMA-fed SS operand, pre-biased code, 16 col2im warps)
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_tc_gpu.py -q -x -s > gpurun_out/r02b_tc.log 2>&1; echo "tc rc=$?"; tail -25 gpurun_out/r02b_tc.log
timeout -s KILL 300 python -m pytest tests/test_sharded_gpu.py tests/test_parity_gpu.py -q -s > gpurun_out/r02b_sh.log 2>&1; echo "sharded rc=$?"; tail -12 gpurun_out/r02b_sh.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout -s KILL 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_cfg2.json 2> gpurun_out/r02b_bench_cfg2.err; echo "rc=$?"; tail -c 1800 gpurun_out/r02b_bench_cfg2.json; tail -5 gpurun_out/r02b_bench_cfg2.err
timeout -s KILL 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02b_bench_cfg5.json 2> gpurun_out/r02b_bench_cfg5.err; echo "rc=$?"; tail -c 1800 gpurun_out/r02b_bench_cfg5.json; tail -5 gpurun_out/r02b_bench_cfg5.err
