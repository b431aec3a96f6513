#!/bin/bash
# Round 2, call c: copy-free blits + fused temporal/a-trous-0 kernel: parity tests, then the bench in the three SVGF modes.
mkdir -p gpurun_out
python -m pytest tests/test_svgf_gpu.py tests/test_ssao_gpu.py tests/test_golden.py tests/test_host_gpu.py tests/test_partition_gpu.py -m gpu -q --maxfail=30 > gpurun_out/r02c_pytest.log 2>&1
tail -12 gpurun_out/r02c_pytest.log
for m in reference alias fused; do
  timeout 600 python bench.py --svgf $m --no-cpu-baseline > gpurun_out/r02c_bench_$m.json 2> gpurun_out/r02c_bench_$m.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02c_bench_$m.json').read().strip().splitlines()[-1])
print('$m', 'ms/step', round(d['ms_per_step'],4), 'svgf', round(d['svgf']['ms_per_frame'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), [ (k['kernel'][:24], round(k['ms']*1e3,1)) for k in d['kernels']])
PY
done
