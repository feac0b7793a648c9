#!/bin/bash
# Round 2, call 30 (one B200): video synthesis flush with fixed thread roles (column walk) - tests, stand-alone kernel time, benches
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_tc_gpu.py tests/test_sharded_gpu.py tests/test_zz_embed3d_gpu.py tests/test_input_pipeline_gpu.py -q -x 2>&1 | tail -2
CDL_SYN_SWEEP=1 timeout -s KILL 300 python -m pytest tests/test_tc_gpu.py -q -x 2>&1 | tail -1
timeout -s KILL 200 python scripts/syn_phase.py 4 0 128 2>&1 | tail -1 | tee gpurun_out/r02ai_phase.json
timeout -s KILL 200 python scripts/syn_phase.py 16 0 2>&1 | tail -1 | tee gpurun_out/r02ai_phase16.json
timeout -s KILL 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02ai_bench_cfg2.json 2> gpurun_out/r02ai_bench_cfg2.err; echo "cfg2 rc=$?"
timeout -s KILL 500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02ai_bench_cfg5.json 2> gpurun_out/r02ai_bench_cfg5.err; echo "cfg5 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02ai_bench_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d["roofline"]
    print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), {k:round(v["avg_launch_ms"],4) for k,v in r["kernels"].items()}, "periter", round(r["per_iteration"]["frac"],3), d["clocks"]["sm_mhz"])
PY
