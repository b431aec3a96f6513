#!/bin/bash
# axis-sorted child slots + near-side-first pop order (VHR_CHILD_SORT=2) against the collapse order (0)
mkdir -p gpurun_out
L=gpurun_out/r01j_trace.log
rm -f $L
for cfg in "0 3000000" "2 3000000" "0 260000" "2 260000"; do
  set -- $cfg
  echo "== child sort $1 tris $2" >> $L
  VHR_CHILD_SORT=$1 VHR_RAYGEN_VARIANT=0 timeout 300 python tools/time_trace.py $2 1920 1080 10 >> $L 2>&1
done
grep "child sort\|update_geometry\|gbuffer\|shadow only\|ao 1spp\|shadow+ao1\|reflection only\|reference\|rror" $L
VHR_CHILD_SORT=2 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r01j_pytest_sort2.log 2>&1
tail -3 gpurun_out/r01j_pytest_sort2.log
