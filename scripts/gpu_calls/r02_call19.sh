#!/bin/bash
# Round 2, call 19 (one B200): L2 prefetch policy / distance of the video synthesis kernel and the residual-box prefetch of the
# analysis kernel (per-role cycle counters, 16 clips = 1.4 GB of code), then the default build on the bench
mkdir -p gpurun_out
P=$PWD/cdlnet-video_b200
for v in prof prof_rpf0 prof_rpf3 prof_nohint prof_pf2 prof_pf2nohint; do
  echo "== $v"; CDL_LIB_PATH=$P/libcdl_b200_$v.so timeout -s KILL 200 python scripts/tc_timeline.py 16 2>&1 | grep -v "^$" | tee gpurun_out/r02x_timeline_$v.log
done
timeout -s KILL 300 python -m pytest tests/test_tc_gpu.py -q -x 2>&1 | tail -2
timeout -s KILL 500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02x_bench_cfg5.json 2> gpurun_out/r02x_bench_cfg5.err; echo "cfg5 rc=$?"
timeout -s KILL 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02x_bench_cfg2.json 2> gpurun_out/r02x_bench_cfg2.err; echo "cfg2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02x_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), {k:round(v["avg_launch_ms"],4) for k,v in r["kernels"].items()}, "periter", round(r["per_iteration"]["frac"],3), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "ERR", e)
PY
