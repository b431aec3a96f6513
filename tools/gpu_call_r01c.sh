#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_rt_gpu.py tests/test_textures_gpu.py tests/test_ssr_gpu.py -m gpu -q --maxfail=20 > gpurun_out/r01c_pytest.log 2>&1
tail -4 gpurun_out/r01c_pytest.log
rm -f gpurun_out/r01c_trace.log
for cfg in "0 12" "2 12" "4 4" "4 8" "4 12" "4 16" "4 24" "5 8" "5 16"; do
  set -- $cfg
  echo "== raygen variant $1 batch $2" >> gpurun_out/r01c_trace.log
  VHR_RAYGEN_VARIANT=$1 VHR_LEAF_BATCH=$2 timeout 300 python tools/time_trace.py 3000000 1920 1080 10 >> gpurun_out/r01c_trace.log 2>&1
done
grep "variant\|shadow only\|ao 1spp\|shadow+ao1\|reflection only\|reference" gpurun_out/r01c_trace.log
