#!/bin/bash
# Round 2, call 24 (one B200): 2-D synthesis kernel with incremental tile coordinates, unpredicated interior loads and a column-walking flush
mkdir -p gpurun_out
rm -f gpurun_out/named_config_parity.jsonl
timeout -s KILL 600 python -m pytest tests/test_tc2_gpu.py tests/test_zz_golden_tc2_gpu.py tests/test_named_configs_gpu.py tests/test_parity_gpu.py -q -x -s 2>&1 | grep -E "passed|failed|graph replay|^cfg|Error" | tail -12
TC2_ARMS=tc2,tc2x3 timeout -s KILL 300 python scripts/tc2_bench.py cfg1b cfg4 cfg3 > gpurun_out/r02ac_tc2_bench.jsonl 2> gpurun_out/r02ac_tc2_bench.err; echo "tc2 rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r02ac_tc2_bench.jsonl"):
    d=json.loads(l)
    print(d["config"], {k:round(v,3) for k,v in d.items() if k.endswith("_ms")}, d.get("max_abs_xhat_tc2x3_vs_fp32"))
PY
