#!/bin/bash
# Round 2, call av: how the ray pass responds to tree quality (radix tree vs PLOC, SAH cost printed by the tool) — sizing what a better hierarchy could buy.
mkdir -p gpurun_out
for b in 0 1; do for r in 8 32; do echo "== builder $b radius $r"; VHR_BVH_BUILDER=$b VHR_PLOC_RADIUS=$r VHR_RAYGEN_VARIANT=0 python tools/time_trace.py 3000000 1920 1080 10 2>&1 | grep -E "update_geometry|shadow only|ao 1spp|shadow\+ao1" ; done; done | tee gpurun_out/r02av_trace.log
