#!/bin/bash
# leaf size study (VHR_MAX_LEAF_TRIS = 2 / 3 (default) / 4) + BVH tests on the default and on 4
mkdir -p gpurun_out
L=gpurun_out/r01x_trace.log
rm -f $L
timeout 200 python -m pytest tests/test_rt_gpu.py -m gpu -x -q > gpurun_out/r01x_pytest_default.log 2>&1; tail -1 gpurun_out/r01x_pytest_default.log
VHR_MAX_LEAF_TRIS=4 timeout 200 python -m pytest tests/test_rt_gpu.py -m gpu -x -q > gpurun_out/r01x_pytest_leaf4.log 2>&1; tail -1 gpurun_out/r01x_pytest_leaf4.log
for cfg in "3 3000000" "4 3000000" "2 3000000" "4 260000"; do
  set -- $cfg
  echo "== max leaf tris $1 tris $2" >> $L
  VHR_MAX_LEAF_TRIS=$1 VHR_RAYGEN_VARIANT=0 timeout 200 python tools/time_trace.py $2 1920 1080 10 >> $L 2>&1
done
grep "max leaf\|update_geometry\|gbuffer\|shadow+ao1\|reflection only\|rror" $L
