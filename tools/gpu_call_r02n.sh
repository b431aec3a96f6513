#!/bin/bash
# Round 2, call n: raygen variant 13 (one inlined any-hit traversal for shadow + AO rays) against the default.
mkdir -p gpurun_out
python -m pytest tests/test_rt_gpu.py -m gpu -q --maxfail=30 -k "variants" > gpurun_out/r02n_pytest.log 2>&1; tail -2 gpurun_out/r02n_pytest.log
for v in 0 13 0 13; do echo "== variant $v"; VHR_RAYGEN_VARIANT=$v python tools/time_trace.py 3000000 1920 1080 10 2>&1 | grep -E "shadow only|ao 1spp|ao 2spp|shadow\+ao1|reference" ; done | tee gpurun_out/r02n_trace.log
for v in 0 13; do echo "== 260k variant $v"; VHR_RAYGEN_VARIANT=$v python tools/time_trace.py 260000 1920 1080 10 2>&1 | grep -E "shadow\+ao1|reference" ; done | tee -a gpurun_out/r02n_trace.log
