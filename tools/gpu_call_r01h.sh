#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r01h_trace.log
for r in 8 32; do
  echo "== ploc radius $r" >> gpurun_out/r01h_trace.log
  VHR_BVH_BUILDER=1 VHR_PLOC_RADIUS=$r VHR_RAYGEN_VARIANT=0 timeout 300 python tools/time_trace.py 3000000 1920 1080 10 >> gpurun_out/r01h_trace.log 2>&1
  VHR_BVH_BUILDER=1 VHR_PLOC_RADIUS=$r VHR_RAYGEN_VARIANT=0 timeout 300 python tools/time_trace.py 260000 1920 1080 10 >> gpurun_out/r01h_trace.log 2>&1
done
echo "== ploc radius 16 + BVH_CT 0.7 / 1.5" >> gpurun_out/r01h_trace.log
VHR_BVH_CT=0.7 VHR_BVH_BUILDER=1 VHR_RAYGEN_VARIANT=0 timeout 300 python tools/time_trace.py 3000000 1920 1080 10 >> gpurun_out/r01h_trace.log 2>&1
VHR_BVH_CT=1.5 VHR_BVH_BUILDER=1 VHR_RAYGEN_VARIANT=0 timeout 300 python tools/time_trace.py 3000000 1920 1080 10 >> gpurun_out/r01h_trace.log 2>&1
grep "radius\|update_geometry\|gbuffer\|shadow only\|ao 1spp\|shadow+ao1\|reflection only\|reference\|rror" gpurun_out/r01h_trace.log
