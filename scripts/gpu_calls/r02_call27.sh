#!/bin/bash
# Round 2, call 17 (8 GPUs): the driver's scaling command on the temporally sharded 1080p clip at N = 8 and N = 4
mkdir -p gpurun_out
for n in 8; do
  timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2970$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r02af_bench_n$n.json 2> gpurun_out/r02af_bench_n$n.err; echo "bench n=$n rc=$?"; tail -3 gpurun_out/r02af_bench_n$n.err
done
python - <<'PY'
import json
for n in (8,):
    try:
        d=json.loads(open(f"gpurun_out/r02af_bench_n{n}.json").read().strip().splitlines()[-1])
        r=d["roofline"]
        print(n, "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "check", d["config"]["sharded_vs_unsharded_max_abs"], {k:round(v["avg_launch_ms"],3) for k,v in r["kernels"].items()}, "xch", r["exchange_ms"], d["clocks"]["sm_mhz"])
    except Exception as e:
        print(n, "ERR", e)
PY
