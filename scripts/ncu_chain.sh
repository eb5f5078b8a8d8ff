#!/bin/bash
# ncu --set full capture of the mask chain (temporal + act + dst kernels) of ONE batch of a configuration, third batch
# of scripts/chain_run.py; leaves the raw CSV page in gpurun_out/ (the .ncu-rep is dropped unless KEEP_REP=1).
# usage: scripts/ncu_chain.sh <tag> W H n B [dy] [mask]
tag=$1; shift
out=gpurun_out/chain_$tag
# launches of the chain per batch: temporal, act, dst_sparse, dst_dense -> skip the first two batches
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'temporal3_kernel|temporal2_kernel|act4_kernel|act_kernel|dst_sparse_kernel|dst_dense_kernel' -s 8 -c 4 \
  -o $out -f python scripts/chain_run.py "$@" > $out.log 2>&1
ncu -i $out.ncu-rep --page raw --csv > $out.raw.csv 2>/dev/null
if [ "$SOURCE_PAGE" = "1" ]; then ncu -i $out.ncu-rep --page source --csv > $out.source.csv 2>/dev/null; fi
ls -la $out.ncu-rep; [ "$KEEP_REP" = "1" ] || rm -f $out.ncu-rep
tail -4 $out.log
