#!/bin/bash
# Round 2, call 10 (one B200): where the video synthesis kernel's tile time goes (per-role cycle counters, phase ceilings,
# TMEM-read-only / no-TMEM-read builds), the gated embed3d tests, ncu --set full of the video kernels and the 2-D kernels.
mkdir -p gpurun_out
P=$PWD/cdlnet-video_b200
echo "== timeline"; CDL_LIB_PATH=$P/libcdl_b200_prof.so timeout -s KILL 120 python scripts/tc_timeline.py 4 > gpurun_out/r02o_timeline.log 2>&1; echo "rc=$?"; cat gpurun_out/r02o_timeline.log
echo "== phases"; timeout -s KILL 300 python scripts/syn_phase.py 4 0 64 128 192 256 448 > gpurun_out/r02o_phase.json 2> gpurun_out/r02o_phase.err; echo "rc=$?"; cat gpurun_out/r02o_phase.json; tail -3 gpurun_out/r02o_phase.err
for e in 1 2; do echo "== CDL_SYN_EXP=$e"; CDL_LIB_PATH=$P/libcdl_b200_exp$e.so timeout -s KILL 200 python scripts/syn_phase.py 4 0 128 2>&1 | tail -1 | tee gpurun_out/r02o_exp$e.json; done
echo "== embed3d"; CDL_RUN_EXPERIMENTAL=1 timeout -s KILL 300 python -m pytest tests/test_zz_embed3d_gpu.py -q -s > gpurun_out/r02o_embed3d.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r02o_embed3d.log
echo "== ncu video"; bash scripts/gpu_ncu.sh r02o_ncu3d 2>&1 | tail -4
echo "== ncu 2-D"; bash scripts/gpu_ncu_tc2.sh r02o_ncu2d cfg4 tc2 2>&1 | tail -5
ls -la gpurun_out | tail -20
