#!/bin/bash
# Round 2, re-entry call: the container was re-created and gpurun_out/ of the earlier round-2 calls is gone.
# Full GPU suite + smoke + benches of the round-2 kernels on one B200; everything is written under gpurun_out/r02n_*.
mkdir -p gpurun_out
rm -f gpurun_out/named_config_parity.jsonl
timeout -s KILL 1200 python -m pytest tests -m gpu -q -rs --durations=15 > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/r02n_pytest.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02n_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02n_smoke.log
timeout -s KILL 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02n_bench_cfg2.json 2> gpurun_out/r02n_bench_cfg2.err; echo "cfg2 rc=$?"; tail -3 gpurun_out/r02n_bench_cfg2.err
timeout -s KILL 500 python bench.py --steps 3 --warmup 3 > gpurun_out/r02n_bench_cfg5.json 2> gpurun_out/r02n_bench_cfg5.err; echo "cfg5 rc=$?"; tail -3 gpurun_out/r02n_bench_cfg5.err
python - <<'PY'
import json
for f in ("gpurun_out/r02n_bench_cfg2.json","gpurun_out/r02n_bench_cfg5.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", d.get("e2e",{}).get("value"), "roof", r.get("kernel"), round(r["frac"],3), {k:round(v["avg_launch_ms"],4) for k,v in r["kernels"].items()}, "periter", r.get("per_iteration"))
    except Exception as e:
        print(f, "ERR", e)
PY
TC2_ARMS=fp32,tc2,tc2x3 timeout -s KILL 300 python scripts/tc2_bench.py cfg1b cfg4 cfg3 > gpurun_out/r02n_tc2_bench.jsonl 2> gpurun_out/r02n_tc2_bench.err; echo "tc2 rc=$?"; tail -6 gpurun_out/r02n_tc2_bench.jsonl; tail -3 gpurun_out/r02n_tc2_bench.err
cat gpurun_out/named_config_parity.jsonl 2>/dev/null | tail -20
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
