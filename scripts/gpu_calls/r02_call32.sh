#!/bin/bash
# Round 2, call 32 (one B200): last full GPU suite + smoke on the final tree
mkdir -p gpurun_out
rm -f gpurun_out/named_config_parity.jsonl
timeout -s KILL 1200 python -m pytest tests -m gpu -q -rs > gpurun_out/r02ak_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|SKIPPED|^FAILED" gpurun_out/r02ak_pytest.log | tail -6
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ak_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02ak_smoke.log
timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02ak_bench.json 2>/dev/null; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02ak_bench.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"
