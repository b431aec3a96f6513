#!/bin/bash
# Round 2, call a: compute-sanitizer evidence on the smoke frame and the in-process fused partition test.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
SMOKE='import __graft_entry__ as g; g.smoke()'
timeout 600 $CS --tool memcheck --leak-check no --print-limit 20 python -c "$SMOKE" > gpurun_out/r02a_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"
timeout 900 $CS --tool racecheck --racecheck-report all --print-limit 20 python -c "$SMOKE" > gpurun_out/r02a_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"
timeout 600 $CS --tool initcheck --print-limit 20 python -c "$SMOKE" > gpurun_out/r02a_initcheck_smoke.log 2>&1; echo "initcheck smoke rc=$?"
timeout 600 $CS --tool synccheck --print-limit 20 python -c "$SMOKE" > gpurun_out/r02a_synccheck_smoke.log 2>&1; echo "synccheck smoke rc=$?"
timeout 900 $CS --tool memcheck --leak-check no --print-limit 20 python -m pytest tests/test_partition_gpu.py -x -q -k "2" > gpurun_out/r02a_memcheck_partition.log 2>&1; echo "memcheck partition rc=$?"
timeout 900 $CS --tool racecheck --racecheck-report all --print-limit 20 python -m pytest tests/test_partition_gpu.py -x -q -k "2" > gpurun_out/r02a_racecheck_partition.log 2>&1; echo "racecheck partition rc=$?"
timeout 900 $CS --tool memcheck --leak-check no --print-limit 20 python -m pytest tests/test_svgf_gpu.py tests/test_ssao_gpu.py tests/test_ssr_gpu.py tests/test_composition_gpu.py -x -q > gpurun_out/r02a_memcheck_passes.log 2>&1; echo "memcheck passes rc=$?"
for f in gpurun_out/r02a_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|\[smoke\] ok" $f | tail -4; done
