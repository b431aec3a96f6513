#!/bin/bash
# cost-optimal collapse (VHR_COLLAPSE=1) against the greedy one: BVH tests under it, fill rate and ray-pass times
mkdir -p gpurun_out
L=gpurun_out/r01v_trace.log
rm -f $L
VHR_COLLAPSE=1 timeout 300 python -m pytest tests/test_rt_gpu.py -m gpu -x -q 2>&1 | tail -3
for cfg in "0 1.0 3000000" "1 1.0 3000000" "1 2.0 3000000" "1 0.5 3000000" "1 1.0 260000"; do
  set -- $cfg
  echo "== collapse $1 node cost $2 tris $3" >> $L
  VHR_COLLAPSE=$1 VHR_DP_NODE_COST=$2 VHR_RAYGEN_VARIANT=0 timeout 300 python tools/time_trace.py $3 1920 1080 10 >> $L 2>&1
done
grep "collapse\|update_geometry\|gbuffer\|shadow+ao1\|reflection only\|rror" $L
