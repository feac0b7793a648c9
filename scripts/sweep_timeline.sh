IFS='|' read -ra ALTS <<< "$DEFS"
for D in "${ALTS[@]}"; do
  CDL_TC_PROFILE=1 CDL_NVCC_DEFS="$D" CDL_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
  echo "DEFS [$D]"; python scripts/tc_timeline.py 4 2>&1 | grep -B12 "== analysis" | grep "rank0\|launch"
done
