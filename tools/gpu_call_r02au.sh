#!/bin/bash
# Round 2, call au: compute-sanitizer memcheck over the GPU suite of the head commit (without the 4K / 3 M-triangle configs, which take too long under the tool).
mkdir -p gpurun_out
timeout 2000 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "not config4 and not config5 and not bench_workload and not config3 and not multiprocess" > gpurun_out/r02au_memcheck_suite.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r02au_memcheck_suite.log
