#!/bin/bash
# Round 2, call b: GPU tests against oracle/_ref after the oracle alignment + a baseline bench of the round-1 kernels on this box.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r02b_pytest.log 2>&1
tail -15 gpurun_out/r02b_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1; tail -3 gpurun_out/r02b_smoke.log
timeout 900 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 1500 gpurun_out/r02b_bench.json
python tools/time_svgf.py > gpurun_out/r02b_time_svgf.log 2>&1; cat gpurun_out/r02b_time_svgf.log
