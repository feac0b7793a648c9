for R in 3 4 5 6 7; do
  CDL_NVCC_DEFS="-DCDL_ANA_ROWS=$R" CDL_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
  python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ROWS $R', d['value'], d['ms_per_step'], d['roofline']['per_kernel_ms_per_step'])"
done
