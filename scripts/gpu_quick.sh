#!/bin/bash
# quick GPU iteration: parity tests (+ optional bench)
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -25
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
if [ "$1" = "bench" ]; then timeout 600 python bench.py --steps 5 --warmup 3 ${@:2} | tee gpurun_out/quick_bench.json; fi
