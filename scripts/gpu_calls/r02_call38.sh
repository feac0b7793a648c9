#!/bin/bash
# Round 2, call 38 (one B200): analysis epilogue skips padding-only code chunks after the first iteration (2.3 % of the code traffic at M = 169)
mkdir -p gpurun_out
P=$PWD/cdlnet-video_b200
timeout -s KILL 600 python -m pytest tests/test_tc_gpu.py tests/test_sharded_gpu.py tests/test_zz_embed3d_gpu.py tests/test_input_pipeline_gpu.py -q -x 2>&1 | tail -2
for arm in new prev new prev; do
  lib=$P/libcdl_b200.so; [ $arm = prev ] && lib=$P/libcdl_b200_prev.so
  echo "== $arm"; CDL_LIB_PATH=$lib timeout -s KILL 200 python scripts/syn_phase.py 16 0 2>&1 | tail -1
done | tee gpurun_out/r02aq_phase_ab.txt
for arm in new prev; do
  lib=$P/libcdl_b200.so; [ $arm = prev ] && lib=$P/libcdl_b200_prev.so
  CDL_LIB_PATH=$lib timeout -s KILL 500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$arm cfg5', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), {k:round(v['avg_launch_ms'],3) for k,v in r['kernels'].items()}, d['clocks']['sm_mhz'])"
done | tee gpurun_out/r02aq_bench_ab.txt
