#!/bin/bash
for e in 1 2; do echo "== CDL_SYN_EXP=$e"; CDL_LIB_PATH=$PWD/cdlnet-video_b200/libcdl_b200_exp$e.so timeout -s KILL 200 python scripts/syn_phase.py 4 0 64 2>&1 | tail -1; done
