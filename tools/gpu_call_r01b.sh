#!/bin/bash
# One GPU-box call: the whole -m gpu suite, ray-pass variants at the bench size, a short bench run.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=20 --durations=8 > gpurun_out/r01b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r01b_pytest.log
tail -5 gpurun_out/r01b_pytest.log
for v in 0 2 3; do
  echo "== raygen variant $v" >> gpurun_out/r01b_trace.log
  VHR_RAYGEN_VARIANT=$v timeout 300 python tools/time_trace.py 3000000 1920 1080 10 >> gpurun_out/r01b_trace.log 2>&1
done
grep "variant\|shadow+ao1\|reference\|gbuffer" gpurun_out/r01b_trace.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/r01b_bench.json 2> gpurun_out/r01b_bench.err
tail -c 3000 gpurun_out/r01b_bench.json
