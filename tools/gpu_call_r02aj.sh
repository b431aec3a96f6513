#!/bin/bash
# Round 2, call aj: compute-sanitizer on the kernels changed in the second half of the round (SSAO quad image, SSR quad tap / window estimate, tiled blur, any-hit flag).
mkdir -p gpurun_out
T=gpurun_out/r02aj
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_ssao_gpu.py tests/test_ssr_gpu.py tests/test_raytraced_path_gpu.py -m gpu -q -x > ${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 ${T}_memcheck.log
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 7 python -m pytest tests/test_ssao_gpu.py tests/test_ssr_gpu.py -m gpu -q -x > ${T}_initcheck.log 2>&1; echo "initcheck rc=$?"; tail -4 ${T}_initcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_ssao_gpu.py -m gpu -q -x > ${T}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 ${T}_racecheck.log
