#!/bin/bash
T=cdlnet-video_b200/csrc/selftest/tc_selftest
for cfg in "2 1 64 7 200" "2 1 128 7 200" "2 1 176 7 200" "2 1 256 7 200" "2 0 64 7 200" "2 0 128 7 200" "2 0 176 7 200" "2 0 256 7 200" "1 1 64 7 200" "1 1 176 7 200" "1 0 176 7 200" "2 1 64 7 1 1" "2 0 64 7 1 1"; do
  echo "== $cfg"; timeout 30 $T $cfg 2>&1 | grep -E "timing|probe|PASS|FAIL|error"
done
