#!/bin/bash
# Round 2, call 42 (one B200): sharded driver with zero-embedded filters + the whole GPU suite once more
mkdir -p gpurun_out
rm -f gpurun_out/named_config_parity.jsonl
timeout -s KILL 1200 python -m pytest tests -m gpu -q -rs > gpurun_out/r02au_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|SKIPPED|^FAILED|^E " gpurun_out/r02au_pytest.log | tail -8
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
