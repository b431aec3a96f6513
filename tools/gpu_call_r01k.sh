#!/bin/bash
# any-hit pop order under VHR_CHILD_SORT=2: near side first (VHR_ANY_ORDER=1) or always lowest slot first (0), against the collapse order
mkdir -p gpurun_out
L=gpurun_out/r01k_trace.log
rm -f $L
for cfg in "0 1 3000000" "2 1 3000000" "2 0 3000000" "0 1 260000" "2 1 260000" "2 0 260000" "0 1 1000000" "2 0 1000000"; do
  set -- $cfg
  echo "== child sort $1 any order $2 tris $3" >> $L
  VHR_CHILD_SORT=$1 VHR_ANY_ORDER=$2 VHR_RAYGEN_VARIANT=0 timeout 300 python tools/time_trace.py $3 1920 1080 20 >> $L 2>&1
done
grep "child sort\|shadow only\|ao 1spp\|shadow+ao1\|reference\|rror" $L
