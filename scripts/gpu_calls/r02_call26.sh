#!/bin/bash
# Round 2, call 26 (one B200): final check - full GPU suite, smoke, the driver's two bench arms at N = 1
mkdir -p gpurun_out
rm -f gpurun_out/named_config_parity.jsonl
timeout -s KILL 1200 python -m pytest tests -m gpu -q -rs > gpurun_out/r02ae_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|SKIPPED|^FAILED" gpurun_out/r02ae_pytest.log | tail -6
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ae_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02ae_smoke.log
timeout -s KILL 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02ae_bench_ref.json 2> gpurun_out/r02ae_bench_ref.err; echo "ref rc=$?"
timeout -s KILL 900 python bench.py > gpurun_out/r02ae_bench_default.json 2> gpurun_out/r02ae_bench_default.err; echo "default bench rc=$?"; tail -2 gpurun_out/r02ae_bench_default.err
timeout -s KILL 300 python bench.py --workload cfg2 --steps 10 --warmup 3 > gpurun_out/r02ae_bench_cfg2.json 2> gpurun_out/r02ae_bench_cfg2.err; echo "cfg2 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02ae_bench_default.json","gpurun_out/r02ae_bench_cfg2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "steps", d["steps"], "e2e", round(d["e2e"]["value"],1), "traffic", r["traffic"], "frac", round(r["frac"],3), {k:round(v["avg_launch_ms"],4) for k,v in r["kernels"].items()}, "periter", round(r["per_iteration"]["frac"],3), d["clocks"], "cpu", d.get("cpu_baseline",{}).get("value"), "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "ERR", e)
d=json.loads(open("gpurun_out/r02ae_bench_ref.json").read().strip().splitlines()[-1]); print("ref", d["value"], d["cpu_baseline"]["cores"], d["e2e"])
PY
