#!/bin/bash
# Round 2, call 40 (4 GPUs): the driver's command at N = 4 on the final tree (both arms)
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29704 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/r02as_bench_ref_n4.json 2> gpurun_out/r02as_bench_ref_n4.err; echo "ref rc=$?"
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29705 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r02as_bench_n4.json 2> gpurun_out/r02as_bench_n4.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02as_bench_n4.json").read().strip().splitlines()[-1])
print("n4 value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "check", d["config"]["sharded_vs_unsharded_max_abs"], d["clocks"]["sm_mhz"])
r=json.loads(open("gpurun_out/r02as_bench_ref_n4.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
