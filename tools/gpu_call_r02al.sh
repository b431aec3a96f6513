#!/bin/bash
# Round 2, call al: raygen variants 14 / 15 (first 8 / 12 traversal-stack entries in shared memory) against the default.
mkdir -p gpurun_out
python -m pytest tests/test_rt_gpu.py -m gpu -q --maxfail=30 -k "variants" > gpurun_out/r02al_pytest.log 2>&1; tail -2 gpurun_out/r02al_pytest.log
for v in 0 14 15 0 14 15; do echo "== variant $v"; VHR_RAYGEN_VARIANT=$v python tools/time_trace.py 3000000 1920 1080 10 2>&1 | grep -E "shadow only|ao 1spp|ao 2spp|shadow\+ao1|reference" ; done | tee gpurun_out/r02al_trace.log
