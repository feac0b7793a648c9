#!/bin/bash
# Round 2, call 21 (one B200): ncu traffic captures of the final video kernels (config 2 and config 5 workloads), summarised on the box;
# pipelined host-buffer entry (denoise_host) test + e2e bench
mkdir -p gpurun_out
CMD2="python bench.py --workload cfg2 --steps 1 --warmup 1 --clips 4 --no-cpu-baseline --no-breakdown --no-e2e"
ncu --set full --clock-control none --import-source on -k "regex:^k_tc_(analysis|synthesis)" -s 10 -c 2 -f -o /tmp/ncu_cfg2 $CMD2 > gpurun_out/r02z_ncu_cfg2.log 2>&1; echo "ncu cfg2 exit $?"
python scripts/ncu_summary.py /tmp/ncu_cfg2.ncu-rep > gpurun_out/r02z_ncu_cfg2_summary.txt 2>&1; head -7 gpurun_out/r02z_ncu_cfg2_summary.txt | cut -c1-200
CMD5="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-breakdown --no-e2e"
ncu --set full --clock-control none --import-source on -k "regex:^k_tc_(analysis|synthesis)" -s 10 -c 2 -f -o /tmp/ncu_cfg5 $CMD5 > gpurun_out/r02z_ncu_cfg5.log 2>&1; echo "ncu cfg5 exit $?"
python scripts/ncu_summary.py /tmp/ncu_cfg5.ncu-rep > gpurun_out/r02z_ncu_cfg5_summary.txt 2>&1; head -7 gpurun_out/r02z_ncu_cfg5_summary.txt | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches_cfg2.csv $CMD2 > gpurun_out/r02z_launches_cfg2.log 2>&1; echo "launch list exit $?"
timeout -s KILL 300 python -m pytest tests/test_sharded_gpu.py -q -x 2>&1 | tail -2
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02z_bench_cfg5.json 2> gpurun_out/r02z_bench_cfg5.err; echo "cfg5 rc=$?"; tail -2 gpurun_out/r02z_bench_cfg5.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02z_bench_cfg5.json").read().strip().splitlines()[-1])
print("cfg5 value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", d["e2e"], d["clocks"]["sm_mhz"])
PY
