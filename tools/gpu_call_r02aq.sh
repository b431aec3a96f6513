#!/bin/bash
# Round 2, call aq: cache policy of the triangle-record loads (0 = __ldg, 1 = __ldcs evict-first, 2 = L1::no_allocate).
mkdir -p gpurun_out
for v in 0 1 2 0 1 2; do echo "== tri load mode $v"; if [ $v = 0 ]; then unset VHR_LIB_PATH; else export VHR_LIB_PATH=$PWD/build/ab/libvhr_b200_tri$v.so; fi; python tools/time_trace.py 3000000 1920 1080 10 2>&1 | grep -E "shadow only|ao 1spp|shadow\+ao1|reference" ; done | tee gpurun_out/r02aq_trace.log
