#!/bin/bash
# ncu --set full capture of one temporal3 variant; leaves raw + source CSV pages in gpurun_out/ (the .ncu-rep is dropped)
# usage: scripts/ncu_t3.sh <variant> [W H n B]
v=$1; shift
out=gpurun_out/t3_v$v
T3_ONLY=$v T3_REPS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:temporal3 -s 2 -c 1 -o $out -f python scripts/t3_tune.py "$@" > $out.log 2>&1
ncu -i $out.ncu-rep --page raw --csv > $out.raw.csv 2>/dev/null
ncu -i $out.ncu-rep --page source --csv > $out.source.csv 2>/dev/null
ls -la $out.ncu-rep; rm -f $out.ncu-rep
