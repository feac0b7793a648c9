#!/bin/bash
# Round 2, call 44 (2 GPUs): NCCL user-buffer registration of the exchanged buffers (default) vs none (CDL_NCCL_REGISTER=0)
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_sharded_gpu.py -q 2>&1 | tail -2
run() { tag=$1; shift; env "$@" timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/r02aw_$tag.json 2> gpurun_out/r02aw_$tag.err; echo "$tag rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02aw_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "check", d["config"]["sharded_vs_unsharded_max_abs"], "exchange", d["roofline"]["exchange_ms"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("$tag ERR", e); print(open("gpurun_out/r02aw_$tag.err").read()[-1200:])
PY
}
run reg CDL_DUMMY=1
run noreg CDL_NCCL_REGISTER=0
run reg2 CDL_DUMMY=1
