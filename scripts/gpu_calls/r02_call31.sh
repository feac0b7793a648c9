#!/bin/bash
# Round 2, call 31 (one B200): evict-last L2 policy on the video synthesis kernel's scatter-adds - tests, ncu traffic, A/B bench
mkdir -p gpurun_out
P=$PWD/cdlnet-video_b200
timeout -s KILL 400 python -m pytest tests/test_tc_gpu.py tests/test_sharded_gpu.py -q -x 2>&1 | tail -1
CMD5="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-breakdown --no-e2e"
ncu --set full --clock-control none -k "regex:^k_tc_synthesis" -s 5 -c 1 -f -o /tmp/ncu_hint $CMD5 > gpurun_out/r02aj_ncu_hint.log 2>&1; echo "ncu exit $?"
python scripts/ncu_summary.py /tmp/ncu_hint.ncu-rep 2>&1 | head -6 | cut -c1-150 | tee gpurun_out/r02aj_ncu_hint_summary.txt
for arm in hint nohint hint nohint; do
  lib=$P/libcdl_b200.so; [ $arm = nohint ] && lib=$P/libcdl_b200_nohint.so
  CDL_LIB_PATH=$lib timeout -s KILL 500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$arm', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), {k:round(v['avg_launch_ms'],3) for k,v in r['kernels'].items()}, d['clocks']['sm_mhz'])" | tee -a gpurun_out/r02aj_ab.txt
done
