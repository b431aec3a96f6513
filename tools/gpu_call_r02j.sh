#!/bin/bash
# Round 2, call j: ray-queue kernel (raygen variant 9): parity, timings against variant 0 with a few refill / burst settings; 720p a-trous timing.
mkdir -p gpurun_out
python -m pytest tests/test_rt_gpu.py -m gpu -q --maxfail=30 -k "queue or variants or persistent" > gpurun_out/r02j_pytest.log 2>&1
tail -5 gpurun_out/r02j_pytest.log
for v in 0 9; do echo "== variant $v"; VHR_RAYGEN_VARIANT=$v python tools/time_trace.py 3000000 1920 1080 10 2>&1 | grep -E "shadow|ao |reference" ; done | tee gpurun_out/r02j_trace.log
for cfg in "4 4" "8 2" "8 8" "16 4" "16 8" "24 4" "12 16"; do set -- $cfg; echo "== variant 9 refill $1 burst $2"; VHR_REFILL_IDLE=$1 VHR_BURST=$2 VHR_RAYGEN_VARIANT=9 python tools/time_trace.py 3000000 1920 1080 10 2>&1 | grep -E "shadow only|ao 1spp|shadow\+ao1|reference"; done | tee -a gpurun_out/r02j_trace.log
echo "== 260k"; for v in 0 9; do VHR_RAYGEN_VARIANT=$v python tools/time_trace.py 260000 1920 1080 10 2>&1 | grep -E "shadow\+ao1|reference"; done | tee -a gpurun_out/r02j_trace.log
VHR_TIME_VARIANTS=2 python tools/time_svgf.py 1280 720 2>&1 | tee gpurun_out/r02j_time_svgf_720p.log
