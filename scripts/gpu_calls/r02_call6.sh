#!/bin/bash
# full GPU suite + benches on the round-2 kernels
mkdir -p gpurun_out
rm -f gpurun_out/named_config_parity.jsonl
timeout -s KILL 1200 python -m pytest tests -m gpu -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02j_pytest.log
timeout -s KILL 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout -s KILL 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_bench_cfg2.json 2> gpurun_out/r02j_bench_cfg2.err; echo "rc=$?"; tail -3 gpurun_out/r02j_bench_cfg2.err
timeout -s KILL 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_bench_cfg5.json 2> gpurun_out/r02j_bench_cfg5.err; echo "rc=$?"; tail -3 gpurun_out/r02j_bench_cfg5.err
python - <<'PY'
import json
for f in ("gpurun_out/r02j_bench_cfg2.json","gpurun_out/r02j_bench_cfg5.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "roof", r["kernel"], round(r["frac"],3), {k:round(v["avg_launch_ms"],4) for k,v in r["kernels"].items()}, "periter", round(r["per_iteration"]["frac"],3))
    except Exception as e:
        print(f, "ERR", e)
PY
TC2_ARMS=fp32,tc2,tc2x3 timeout -s KILL 200 python scripts/tc2_bench.py cfg1b cfg4 cfg3 2>&1 | tail -4
