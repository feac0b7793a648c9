#!/bin/bash
# Round 2, call 12 (one B200): the "tap half per CTA" synthesis kernel (CDL_SYN_H=1, default) against the cta_group::2 pair form
mkdir -p gpurun_out
P=$PWD/cdlnet-video_b200
timeout -s KILL 300 python -m pytest tests/test_tc_gpu.py tests/test_sharded_gpu.py -q -x > gpurun_out/r02q_tc.log 2>&1; echo "tc rc=$?"; tail -6 gpurun_out/r02q_tc.log
for h in 1 0; do
  echo "== CDL_SYN_H=$h phases"; CDL_SYN_H=$h timeout -s KILL 300 python scripts/syn_phase.py 4 0 64 128 448 2>&1 | tail -1 | tee gpurun_out/r02q_phase_h$h.json
done
echo "== timeline (half form)"; CDL_LIB_PATH=$P/libcdl_b200_prof.so timeout -s KILL 120 python scripts/tc_timeline.py 4 > gpurun_out/r02q_timeline.log 2>&1; echo "rc=$?"; cat gpurun_out/r02q_timeline.log
for h in 1 0; do
  CDL_SYN_H=$h timeout -s KILL 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02q_bench_cfg2_h$h.json 2> gpurun_out/r02q_bench_cfg2_h$h.err; echo "cfg2 h=$h rc=$?"; tail -2 gpurun_out/r02q_bench_cfg2_h$h.err
done
timeout -s KILL 500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02q_bench_cfg5.json 2> gpurun_out/r02q_bench_cfg5.err; echo "cfg5 rc=$?"; tail -3 gpurun_out/r02q_bench_cfg5.err
python - <<'PY'
import json
for f in ("gpurun_out/r02q_bench_cfg2_h1.json","gpurun_out/r02q_bench_cfg2_h0.json","gpurun_out/r02q_bench_cfg5.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "roof", round(r["frac"],3), {k:round(v["avg_launch_ms"],4) for k,v in r["kernels"].items()}, "periter", round(r["per_iteration"]["frac"],3), d["clocks"], d["e2e"].get("max_abs_diff_vs_device_path"))
    except Exception as e:
        print(f, "ERR", e)
PY
echo "== cfg1 embed timing"; timeout -s KILL 120 python scripts/cfg1_embed_timing.py 2>&1 | tail -1 | tee gpurun_out/r02q_cfg1_embed.json
