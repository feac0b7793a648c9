#!/bin/bash
# Round 2, call 34 (one B200): two MMA-issuing warps per leader CTA in the video kernels (default) vs one (libcdl_b200_mw1.so)
mkdir -p gpurun_out
P=$PWD/cdlnet-video_b200
timeout -s KILL 400 python -m pytest tests/test_tc_gpu.py tests/test_sharded_gpu.py tests/test_zz_embed3d_gpu.py -q -x 2>&1 | tail -2
for arm in mw2 mw1 mw2 mw1; do
  lib=$P/libcdl_b200.so; [ $arm = mw1 ] && lib=$P/libcdl_b200_mw1.so
  echo "== $arm"; CDL_LIB_PATH=$lib timeout -s KILL 200 python scripts/syn_phase.py 4 0 2>&1 | tail -1
  CDL_LIB_PATH=$lib timeout -s KILL 200 python scripts/syn_phase.py 16 0 2>&1 | tail -1
done | tee gpurun_out/r02am_phase_ab.txt
for arm in mw2 mw1; do
  lib=$P/libcdl_b200.so; [ $arm = mw1 ] && lib=$P/libcdl_b200_mw1.so
  CDL_LIB_PATH=$lib timeout -s KILL 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$arm cfg2', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), {k:round(v['avg_launch_ms'],4) for k,v in r['kernels'].items()}, d['clocks']['sm_mhz'])"
  CDL_LIB_PATH=$lib timeout -s KILL 500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$arm cfg5', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), {k:round(v['avg_launch_ms'],3) for k,v in r['kernels'].items()}, d['clocks']['sm_mhz'])"
done | tee gpurun_out/r02am_bench_ab.txt
