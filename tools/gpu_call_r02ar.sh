#!/bin/bash
# Round 2, call ar: the two build-time studies again on the DEFAULT raygen kernel (calls am / aq ran tools/time_trace.py's old default, the persistent variant 1):
# triangle-record load policy (tri1 = __ldcs, tri2 = L1::no_allocate) and 128-byte nodes.
mkdir -p gpurun_out
for v in default tri1 tri2 node128 default tri1 tri2 node128; do echo "== $v"; if [ $v = default ]; then unset VHR_LIB_PATH; else export VHR_LIB_PATH=$PWD/build/ab/libvhr_b200_$v.so; fi; VHR_RAYGEN_VARIANT=0 python tools/time_trace.py 3000000 1920 1080 10 2>&1 | grep -E "shadow only|ao 1spp|shadow\+ao1|reference" ; done | tee gpurun_out/r02ar_trace.log
