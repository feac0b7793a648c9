#!/bin/bash
# Round 2, GPU call 1 (one B200): operand-rounding probe, the whole GPU test suite (incl. the named-config parity tests),
# the three round-1 candidates, the new bench at N = 1 (config 5, and config 2 as the secondary line), ncu of the 2-D kernels.
mkdir -p gpurun_out
T=cdlnet-video_b200/csrc/selftest/tc_selftest
{
  for cfg in "2 0 176 7 1 1" "2 1 176 7 1 1" "1 0 176 7 1 1"; do echo "== probe $cfg"; timeout 30 $T $cfg; echo "rc=$?"; done
} > gpurun_out/r02a_probe.log 2>&1
cat gpurun_out/r02a_probe.log
rm -f gpurun_out/named_config_parity.jsonl
timeout -s KILL 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/r02a_pytest.log | tail -5
grep -E "^(cfg|hot|gdlnet|FAILED|ERROR)" gpurun_out/r02a_pytest.log | head -40
timeout -s KILL 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== candidates"
TC2_CANDIDATES="tc2v2 tc2x3" bash scripts/gpu_tc2.sh 150 150 cfg4 > gpurun_out/r02a_candidates.log 2>&1; tail -30 gpurun_out/r02a_candidates.log
echo "== bench cfg5 N=1"
timeout -s KILL 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02a_bench_cfg5_n1.json 2> gpurun_out/r02a_bench_cfg5_n1.err; echo "rc=$?"; tail -c 3000 gpurun_out/r02a_bench_cfg5_n1.json; tail -5 gpurun_out/r02a_bench_cfg5_n1.err
echo "== bench cfg2 N=1"
timeout -s KILL 300 python bench.py --workload cfg2 --steps 10 --warmup 3 > gpurun_out/r02a_bench_cfg2_n1.json 2> gpurun_out/r02a_bench_cfg2_n1.err; echo "rc=$?"; tail -c 2500 gpurun_out/r02a_bench_cfg2_n1.json; tail -5 gpurun_out/r02a_bench_cfg2_n1.err
echo "== reference arm"
timeout -s KILL 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02a_bench_ref.json 2>&1; tail -c 600 gpurun_out/r02a_bench_ref.json
echo "== ncu 2-D kernels"
bash scripts/gpu_ncu_tc2.sh r02a_ncu2d cfg4 tc2 2>&1 | tail -5
