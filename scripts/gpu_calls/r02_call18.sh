#!/bin/bash
# Round 2, call 18 (4 GPUs): exchange cost of the native slab driver, default NCCL P2P channels vs more channels
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 4 --steps 5 --warmup 3 --no-e2e --no-check > gpurun_out/r02w_$tag.json 2> gpurun_out/r02w_$tag.err; echo "$tag rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02w_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "exchange", d["roofline"]["exchange_ms"], "sum", round(d["roofline"]["exchange_ms_per_step"],2), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$tag ERR", e); print(open("gpurun_out/r02w_$tag.err").read()[-1500:])
PY
}
run default CDL_DUMMY=1
run chan16 NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=16
run chan32 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
run default2 CDL_DUMMY=1
