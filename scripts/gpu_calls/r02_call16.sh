#!/bin/bash
# Round 2, call 16 (one B200): CSR variants on the device, config 1 on the video kernels by default, fused input pipeline;
# ncu traffic of the video kernels with the new tile orders (one launch each at the bench's default workload)
mkdir -p gpurun_out
rm -f gpurun_out/named_config_parity.jsonl
timeout -s KILL 600 python -m pytest tests/test_csr.py tests/test_zz_embed3d_gpu.py tests/test_named_configs_gpu.py tests/test_input_pipeline_gpu.py tests/test_parity_gpu.py -m gpu -q -rs > gpurun_out/r02u_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r02u_tests.log
grep -h "^cfg1" gpurun_out/r02u_tests.log | head -3; head -2 gpurun_out/named_config_parity.jsonl | cut -c1-300
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-breakdown --no-e2e"
ncu --set full --clock-control none --import-source on -k "regex:^k_tc_(analysis|synthesis)" -s 10 -c 2 -f -o gpurun_out/r02u_ncu3d_prof $CMD > gpurun_out/r02u_ncu3d.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/r02u_ncu3d.log
