#!/bin/bash
# Round 2, call k: raygen occupancy variants (10: 48 registers / 10 blocks, 11: 40 / 12) and 4-byte stack entries (12) against the default.
mkdir -p gpurun_out
python -m pytest tests/test_rt_gpu.py -m gpu -q --maxfail=30 -k "queue" > gpurun_out/r02k_pytest.log 2>&1; tail -2 gpurun_out/r02k_pytest.log
for v in 0 10 11 12 0 12; do echo "== variant $v"; VHR_RAYGEN_VARIANT=$v python tools/time_trace.py 3000000 1920 1080 10 2>&1 | grep -E "shadow only|ao 1spp|shadow\+ao1|reflection only|reference" ; done | tee gpurun_out/r02k_trace.log
