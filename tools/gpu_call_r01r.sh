#!/bin/bash
# raygen variant 8 (AO rays sorted by direction inside the CTA) against variant 0
mkdir -p gpurun_out
L=gpurun_out/r01r_trace.log
rm -f $L
timeout 300 python -m pytest tests/test_rt_gpu.py -m gpu -x -q -k "variants" 2>&1 | tail -3
for cfg in "0 3000000" "8 3000000" "0 260000" "8 260000"; do
  set -- $cfg
  echo "== raygen variant $1 tris $2" >> $L
  VHR_RAYGEN_VARIANT=$1 timeout 300 python tools/time_trace.py $2 1920 1080 20 >> $L 2>&1
done
grep "variant\|shadow only\|ao 1spp\|ao 2spp\|shadow+ao1\|reference\|rror" $L
