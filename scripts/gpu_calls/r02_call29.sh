#!/bin/bash
# Round 2, call 25 (one B200): 3-term final D z on the 2-D tensor-core synthesis kernel
mkdir -p gpurun_out
rm -f gpurun_out/named_config_parity.jsonl
timeout -s KILL 600 python -m pytest tests/test_tc2_gpu.py tests/test_zz_golden_tc2_gpu.py tests/test_named_configs_gpu.py tests/test_parity_gpu.py tests/test_input_pipeline_gpu.py -q -x -s 2>&1 | grep -E "passed|failed|^cfg|^hot|^gdlnet|Error|assert" | tail -14
timeout -s KILL 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
TC2_ARMS=tc2,tc2x3 timeout -s KILL 300 python scripts/tc2_bench.py cfg1b cfg4 cfg3 > gpurun_out/r02ah_tc2_bench.jsonl 2> gpurun_out/r02ah_tc2_bench.err; echo "tc2 rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r02ah_tc2_bench.jsonl"):
    d=json.loads(l)
    print(d["config"], {k:round(v,3) for k,v in d.items() if k.endswith("_ms")}, {k:v for k,v in d.items() if k.startswith("max_abs")})
PY
