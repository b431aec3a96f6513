#!/bin/bash
# slot fill rate of the wide nodes (new statistic) for the three scene sizes, both builders
mkdir -p gpurun_out
L=gpurun_out/r01u_trace.log
rm -f $L
for cfg in "1 260000" "1 1000000" "1 3000000" "0 3000000"; do
  set -- $cfg
  echo "== builder $1 tris $2" >> $L
  VHR_BVH_BUILDER=$1 VHR_RAYGEN_VARIANT=0 timeout 300 python tools/time_trace.py $2 1920 1080 10 >> $L 2>&1
done
grep "builder\|update_geometry\|shadow+ao1\|rror" $L
