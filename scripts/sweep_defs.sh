# usage: DEFS="a|b|c" bash scripts/sweep_defs.sh   (each alternative = one nvcc -D string)
IFS='|' read -ra ALTS <<< "$DEFS"
for D in "${ALTS[@]}"; do
  CDL_NVCC_DEFS="$D" CDL_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
  python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('DEFS [$D]', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['per_kernel_ms_per_step'])"
done
