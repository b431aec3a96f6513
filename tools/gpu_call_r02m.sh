#!/bin/bash
# Round 2, call m: a-trous pair kernel skipping the shadow channel's tap work in warps whose whole footprint is settled: parity, A/B timings.
mkdir -p gpurun_out
python -m pytest tests/test_svgf_gpu.py tests/test_golden.py tests/test_baseline_configs_gpu.py tests/test_partition_gpu.py tests/test_host_gpu.py -m gpu -q --maxfail=30 > gpurun_out/r02m_pytest.log 2>&1
tail -4 gpurun_out/r02m_pytest.log
for sk in 0 1; do
  echo "== VHR_ATROUS_SKIP_SETTLED=$sk (noise input: nothing to skip)"; VHR_ATROUS_SKIP_SETTLED=$sk VHR_TIME_VARIANTS=2 python tools/time_svgf.py 2>&1
  VHR_ATROUS_SKIP_SETTLED=$sk timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02m_bench_skip$sk.json 2> gpurun_out/r02m_bench_skip$sk.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02m_bench_skip$sk.json').read().strip().splitlines()[-1])
print('skip $sk: ms/step', round(d['ms_per_step'],4), 'svgf', round(d['svgf']['ms_per_frame'],4), [ (k['kernel'][:28], round(k['ms']*1e3,1)) for k in d['kernels']])
PY
done 2>&1 | tee gpurun_out/r02m_ab.log
