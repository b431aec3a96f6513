#!/bin/bash
# PLOC (deterministic numbering) as the builder: full GPU suite under it + ray-pass timings against the radix tree
mkdir -p gpurun_out
L=gpurun_out/r01i_trace.log
rm -f $L
VHR_BVH_BUILDER=1 VHR_PLOC_RADIUS=8 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r01i_pytest_ploc.log 2>&1
tail -3 gpurun_out/r01i_pytest_ploc.log
for cfg in "0 8 260000" "1 8 260000" "0 8 1000000" "1 8 1000000" "1 4 3000000" "1 8 3000000" "1 8 3000000"; do
  set -- $cfg
  echo "== builder $1 radius $2 tris $3" >> $L
  VHR_BVH_BUILDER=$1 VHR_PLOC_RADIUS=$2 VHR_RAYGEN_VARIANT=0 timeout 300 python tools/time_trace.py $3 1920 1080 10 >> $L 2>&1
done
grep "builder\|update_geometry\|gbuffer\|shadow+ao1\|reflection only\|reference\|rror" $L
