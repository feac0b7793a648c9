#!/bin/bash
# Round 2, call 11 (2 GPUs): NCCL slab test + the driver's bench command on the temporally sharded clip at N = 2
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_sharded_gpu.py -q -s -rs > gpurun_out/r02p_sharded.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02p_sharded.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02p_bench_n2.json 2> gpurun_out/r02p_bench_n2.err; echo "bench rc=$?"; tail -c 3500 gpurun_out/r02p_bench_n2.json; tail -8 gpurun_out/r02p_bench_n2.err
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02p_bench_ref_n2.json 2> gpurun_out/r02p_bench_ref_n2.err; echo "ref rc=$?"; tail -c 1200 gpurun_out/r02p_bench_ref_n2.json; tail -3 gpurun_out/r02p_bench_ref_n2.err
