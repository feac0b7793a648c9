#!/bin/bash
# 2 GPUs: exchange cost of the native slab driver, default NCCL P2P channels vs tuned
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --no-check > gpurun_out/r02m_$tag.json 2> gpurun_out/r02m_$tag.err; echo "$tag rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02m_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "exchange", d["roofline"]["exchange_ms"], "sum", round(d["roofline"]["exchange_ms_per_step"],2))
except Exception as e:
    print("$tag ERR", e); print(open("gpurun_out/r02m_$tag.err").read()[-1500:])
PY
}
run default CDL_DUMMY=1
run chan32 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
run debug NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,P2P
grep -i "p2p\|channel\|NVLS\|nchannels" gpurun_out/r02m_debug.err | head -30
