#!/bin/bash
# Round 2, call 15 (one B200): frame-synchronous tile order of the video synthesis kernel (CDL_SYN_SWEEP) + evict-first code loads
mkdir -p gpurun_out
CDL_SYN_SWEEP=1 timeout -s KILL 300 python -m pytest tests/test_tc_gpu.py tests/test_sharded_gpu.py -q -x > gpurun_out/r02t_tc_sweep1.log 2>&1; echo "tc sweep=1 rc=$?"; tail -3 gpurun_out/r02t_tc_sweep1.log
timeout -s KILL 300 python -m pytest tests/test_tc_gpu.py tests/test_input_pipeline_gpu.py -q -x > gpurun_out/r02t_tc.log 2>&1; echo "tc rc=$?"; tail -3 gpurun_out/r02t_tc.log
for sw in 1 0; do
  CDL_SYN_SWEEP=$sw timeout -s KILL 500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02t_bench_cfg5_sw$sw.json 2> gpurun_out/r02t_bench_cfg5_sw$sw.err; echo "cfg5 sweep=$sw rc=$?"; tail -2 gpurun_out/r02t_bench_cfg5_sw$sw.err
  CDL_SYN_SWEEP=$sw timeout -s KILL 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02t_bench_cfg2_sw$sw.json 2> gpurun_out/r02t_bench_cfg2_sw$sw.err; echo "cfg2 sweep=$sw rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02t_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), {k:round(v["avg_launch_ms"],4) for k,v in r["kernels"].items()}, "periter", round(r["per_iteration"]["frac"],3), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "ERR", e)
PY
