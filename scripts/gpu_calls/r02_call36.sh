#!/bin/bash
# Round 2, call 36 (one B200): A ring of the video synthesis kernel as 4 slots x 3 K-steps (12 KB) vs 3 slots x 4 K-steps (16 KB), same 48 KB
mkdir -p gpurun_out
P=$PWD/cdlnet-video_b200
CDL_LIB_PATH=$P/libcdl_b200_k3s4.so timeout -s KILL 400 python -m pytest tests/test_tc_gpu.py tests/test_sharded_gpu.py -q -x 2>&1 | tail -2
for arm in k3s4 k4s3 k3s4 k4s3; do
  lib=$P/libcdl_b200.so; [ $arm = k3s4 ] && lib=$P/libcdl_b200_k3s4.so
  echo "== $arm"; CDL_LIB_PATH=$lib timeout -s KILL 200 python scripts/syn_phase.py 4 0 2>&1 | tail -1
  CDL_LIB_PATH=$lib timeout -s KILL 200 python scripts/syn_phase.py 16 0 2>&1 | tail -1
done | tee gpurun_out/r02ao_phase_ab.txt
for arm in k3s4 k4s3; do
  lib=$P/libcdl_b200.so; [ $arm = k3s4 ] && lib=$P/libcdl_b200_k3s4.so
  CDL_LIB_PATH=$lib timeout -s KILL 500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$arm cfg5', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), {k:round(v['avg_launch_ms'],3) for k,v in r['kernels'].items()}, d['clocks']['sm_mhz'])"
done | tee gpurun_out/r02ao_bench_ab.txt
