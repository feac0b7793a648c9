#!/bin/bash
# Round 2, call 37 (one B200): smaller odd filter extents zero-embedded into the 7x7x7 tensor-core kernels; full suite again
mkdir -p gpurun_out
rm -f gpurun_out/named_config_parity.jsonl
timeout -s KILL 1200 python -m pytest tests -m gpu -q -rs > gpurun_out/r02ap_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|SKIPPED|^FAILED|^E " gpurun_out/r02ap_pytest.log | tail -8
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
