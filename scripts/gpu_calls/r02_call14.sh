#!/bin/bash
# Round 2, call 14 (one B200): analysis tile order (frames inside an h-band), fused input pipeline tests, full GPU suite, benches
mkdir -p gpurun_out
rm -f gpurun_out/named_config_parity.jsonl
timeout -s KILL 1200 python -m pytest tests -m gpu -q -rs > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02s_pytest.log
timeout -s KILL 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02s_bench_cfg2.json 2> gpurun_out/r02s_bench_cfg2.err; echo "cfg2 rc=$?"; tail -2 gpurun_out/r02s_bench_cfg2.err
timeout -s KILL 500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02s_bench_cfg5.json 2> gpurun_out/r02s_bench_cfg5.err; echo "cfg5 rc=$?"; tail -3 gpurun_out/r02s_bench_cfg5.err
python - <<'PY'
import json
for f in ("gpurun_out/r02s_bench_cfg2.json","gpurun_out/r02s_bench_cfg5.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "roof", round(r["frac"],3), {k:round(v["avg_launch_ms"],4) for k,v in r["kernels"].items()}, "periter", round(r["per_iteration"]["frac"],3), d["clocks"])
    except Exception as e:
        print(f, "ERR", e)
PY
