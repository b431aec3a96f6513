#!/bin/bash
# raygen CTA size: four-warp (variant 0), one-warp (6), two-warp (7) CTAs
mkdir -p gpurun_out
L=gpurun_out/r01m_trace.log
rm -f $L
timeout 300 python -m pytest tests/test_rt_gpu.py -m gpu -x -q -k "variants" 2>&1 | tail -2
for cfg in "0 3000000" "6 3000000" "7 3000000" "0 260000" "6 260000" "7 260000"; do
  set -- $cfg
  echo "== raygen variant $1 tris $2" >> $L
  VHR_RAYGEN_VARIANT=$1 timeout 300 python tools/time_trace.py $2 1920 1080 20 >> $L 2>&1
done
grep "variant\|shadow only\|ao 1spp\|shadow+ao1\|reflection only\|reference\|rror" $L
