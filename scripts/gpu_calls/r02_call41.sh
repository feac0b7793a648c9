#!/bin/bash
# Round 2, call 41 (8 GPUs): the scaling curve N = 1, 2, 4, 8 back to back on ONE box (same GPUs, same power envelope), final tree
mkdir -p gpurun_out
timeout -s KILL 400 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02at_scale_n1.json 2> gpurun_out/r02at_scale_n1.err; echo "n=1 rc=$?"
for n in 2 4 8; do
  timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r02at_scale_n$n.json 2> gpurun_out/r02at_scale_n$n.err; echo "n=$n rc=$?"
done
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.loads(open(f"gpurun_out/r02at_scale_n{n}.json").read().strip().splitlines()[-1])
        if n==1: base=d["value"]
        print(n, "value", round(d["value"],1), "x", round(d["value"]/base,2), "eff", round(d["value"]/base/n,3), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "clk", d["clocks"]["sm_mhz"], "xch", d["roofline"].get("exchange_ms",{}).get("median") if n>1 else None)
    except Exception as e:
        print(n, "ERR", e)
PY
