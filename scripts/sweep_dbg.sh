for M in ${MODES:-0 2}; do
  CDL_TC_DBG_MODE=$M python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MODE $M', d['value'], d['ms_per_step'], d['roofline']['per_kernel_ms_per_step'])"
done
