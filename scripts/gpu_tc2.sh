#!/bin/bash
# The 2-D tensor-core kernels on the B200 box: parity tests of the default kernels, then - each in its OWN process, a
# trapped kernel poisons the CUDA context - the gated tests of the two candidates, the smoke, per-config timing.
#   bash scripts/gpu_tc2.sh [test-timeout] [bench-timeout] [configs...]      TC2_ARMS (default fp32,tc2) selects the arms of
#   the main timing run, TC2_CANDIDATES (default "tc2v2 tc2x3") the candidate arms timed afterwards in separate processes
mkdir -p gpurun_out
L=gpurun_out/tc2_bringup.log
{
  echo "== default kernels"
  timeout -s KILL ${1:-150} python -m pytest tests/test_tc2_gpu.py tests/test_zz_golden_tc2_gpu.py -q -s 2>&1 | tail -40
  echo "== candidate: 3-term analysis (CDL_TC2D_ANA=3)"
  CDL_RUN_EXPERIMENTAL=1 timeout -s KILL ${1:-150} python -m pytest tests/test_tc2_gpu.py -q -s -k "x3" 2>&1 | tail -40
  echo "== candidate: write-once col2im (CDL_TC2D_SYN=2)"
  CDL_RUN_EXPERIMENTAL=1 timeout -s KILL ${1:-150} python -m pytest tests/test_tc2_gpu.py -q -s -k "v2" 2>&1 | tail -40
  echo "== candidate: config 1 on the video kernels (CDL_EMBED3D=1)"
  CDL_RUN_EXPERIMENTAL=1 timeout -s KILL ${1:-150} python -m pytest tests/test_zz_embed3d_gpu.py -q -s 2>&1 | tail -20
} > $L 2>&1
cat $L
timeout -s KILL 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
T=${2:-120}; shift; shift
TC2_ARMS=${TC2_ARMS:-fp32,tc2} timeout -s KILL $T python scripts/tc2_bench.py ${@:-cfg1b cfg4 cfg3} 2>&1 | tail -20
# the candidates, each in its own process (against the default tensor-core arm, config 4 only): set TC2_CANDIDATES="" to skip
for cand in ${TC2_CANDIDATES-tc2v2 tc2x3}; do
  echo "== candidate arm $cand"
  TC2_ARMS=tc2,$cand timeout -s KILL 60 python scripts/tc2_bench.py cfg4 2>&1 | tail -3
done
