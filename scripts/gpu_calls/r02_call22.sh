#!/bin/bash
# Round 2, call 22 (2 GPUs): NCCL test + the driver's command at N = 2 with the pipelined host-buffer entry; launch list of our kernels (1 GPU)
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_sharded_gpu.py -q 2>&1 | tail -2
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02aa_bench_n2.json 2> gpurun_out/r02aa_bench_n2.err; echo "bench rc=$?"; tail -3 gpurun_out/r02aa_bench_n2.err | grep -v "^\*\|OMP"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02aa_bench_n2.json").read().strip().splitlines()[-1])
print("n2 value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", d["e2e"], "check", d["config"]["sharded_vs_unsharded_max_abs"], d["clocks"]["sm_mhz"])
PY
CMD2="python bench.py --workload cfg2 --steps 1 --warmup 1 --clips 4 --no-cpu-baseline --no-breakdown --no-e2e"
CUDA_VISIBLE_DEVICES=0 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_" -c 400 --csv --log-file gpurun_out/r02aa_launches_cfg2.csv $CMD2 > gpurun_out/r02aa_launches_cfg2.log 2>&1; echo "launch list exit $?"
