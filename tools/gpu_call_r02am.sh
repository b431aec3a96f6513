#!/bin/bash
# Round 2, call am: 8-wide nodes padded from 80 to 128 bytes (one cache line each; build/ab/libvhr_b200_node128.so) against the default.
mkdir -p gpurun_out
VHR_LIB_PATH=$PWD/build/ab/libvhr_b200_node128.so python -m pytest tests/test_rt_gpu.py -m gpu -q --maxfail=5 -k "not variants" > gpurun_out/r02am_pytest.log 2>&1; tail -2 gpurun_out/r02am_pytest.log
for v in default node128 default node128; do echo "== $v"; if [ $v = node128 ]; then export VHR_LIB_PATH=$PWD/build/ab/libvhr_b200_node128.so; else unset VHR_LIB_PATH; fi; python tools/time_trace.py 3000000 1920 1080 10 2>&1 | grep -E "shadow only|ao 1spp|shadow\+ao1|reference" ; done | tee gpurun_out/r02am_trace.log
