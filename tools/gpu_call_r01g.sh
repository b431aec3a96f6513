#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r01g_trace.log
for cfg in "1 0" "1 1" "0 1"; do
  set -- $cfg
  echo "== morton mode $1 child sort $2" >> gpurun_out/r01g_trace.log
  VHR_MORTON_MODE=$1 VHR_CHILD_SORT=$2 VHR_RAYGEN_VARIANT=0 timeout 300 python tools/time_trace.py 3000000 1920 1080 10 >> gpurun_out/r01g_trace.log 2>&1
done
grep "mode\|update_geometry\|gbuffer\|shadow only\|ao 1spp\|shadow+ao1\|reflection only\|reference" gpurun_out/r01g_trace.log
