#!/bin/bash
# Round 2, call 20 (one B200): full GPU suite + smoke on the final kernels, ncu traffic captures (config 2 and config 5 workloads),
# final 1-GPU bench lines
mkdir -p gpurun_out
rm -f gpurun_out/named_config_parity.jsonl
timeout -s KILL 1200 python -m pytest tests -m gpu -q -rs -s > gpurun_out/r02y_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|SKIPPED|^FAILED|graph replay" gpurun_out/r02y_pytest.log | tail -8
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02y_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02y_smoke.log
CMD2="python bench.py --workload cfg2 --steps 1 --warmup 1 --clips 4 --no-cpu-baseline --no-breakdown --no-e2e"
ncu --set full --clock-control none --import-source on -k "regex:^k_tc_(analysis|synthesis)" -s 10 -c 2 -f -o gpurun_out/r02y_ncu_cfg2 $CMD2 > gpurun_out/r02y_ncu_cfg2.log 2>&1; echo "ncu cfg2 exit $?"
CMD5="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-breakdown --no-e2e"
ncu --set full --clock-control none --import-source on -k "regex:^k_tc_(analysis|synthesis)" -s 10 -c 2 -f -o gpurun_out/r02y_ncu_cfg5 $CMD5 > gpurun_out/r02y_ncu_cfg5.log 2>&1; echo "ncu cfg5 exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02y_launches_cfg2.csv $CMD2 > gpurun_out/r02y_launches_cfg2.log 2>&1; echo "launch list exit $?"
timeout -s KILL 300 python bench.py --workload cfg2 --steps 10 --warmup 3 > gpurun_out/r02y_bench_cfg2.json 2> gpurun_out/r02y_bench_cfg2.err; echo "cfg2 rc=$?"
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02y_bench_cfg5.json 2> gpurun_out/r02y_bench_cfg5.err; echo "cfg5 rc=$?"; tail -2 gpurun_out/r02y_bench_cfg5.err
timeout -s KILL 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02y_bench_ref.json 2> gpurun_out/r02y_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json,glob
for f in ("gpurun_out/r02y_bench_cfg2.json","gpurun_out/r02y_bench_cfg5.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "traffic", r["traffic"], {k:round(v["avg_launch_ms"],4) for k,v in r["kernels"].items()}, "periter", round(r["per_iteration"]["frac"],3), d["clocks"]["sm_mhz"], d.get("cpu_baseline",{}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
print(open("gpurun_out/r02y_bench_ref.json").read()[-700:])
PY
