#!/bin/bash
# 2 GPUs: NCCL slab test + bench on the temporally sharded clip (native driver, cdl_forward_sharded)
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_sharded_gpu.py tests/test_zz_golden_tc2_gpu.py -q -s > gpurun_out/r02l_sharded.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02l_sharded.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02l_bench_n2.json 2> gpurun_out/r02l_bench_n2.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/r02l_bench_n2.json; tail -8 gpurun_out/r02l_bench_n2.err
