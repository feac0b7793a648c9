for D in "-DCDL_SYN_CONVERGED=1" "-DNOTHING=1"; do
  CDL_NVCC_DEFS="$D" CDL_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
  for M in 256 768 0; do
  CDL_TC_DBG_MODE=$M python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('COMBO [$D] MODE $M', round(d['ms_per_step'],3), d['roofline']['per_kernel_ms_per_step'])"
  done
done
