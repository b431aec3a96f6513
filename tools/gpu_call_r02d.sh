#!/bin/bash
# Round 2, call d: a-trous pair kernel with staged variance rows (S > 1), 32-bit id compares, quadratic normal-weight polynomial; RY=3 experiment.
mkdir -p gpurun_out
python -m pytest tests/test_svgf_gpu.py tests/test_golden.py tests/test_baseline_configs_gpu.py -m gpu -q --maxfail=30 > gpurun_out/r02d_pytest.log 2>&1
tail -8 gpurun_out/r02d_pytest.log
grep -h "parity\] atrous v2 step" gpurun_out/r02d_pytest.log | head -5
echo "== TR=4 (default)"; VHR_TIME_VARIANTS=2 python tools/time_svgf.py 2>&1 | tee gpurun_out/r02d_time_tr4.log
echo "== TR=3 RY=3";      VHR_ATROUS_TR=3 VHR_TIME_VARIANTS=2 python tools/time_svgf.py 2>&1 | tee gpurun_out/r02d_time_tr3.log
timeout 600 python bench.py --svgf alias --no-cpu-baseline > gpurun_out/r02d_bench_alias.json 2> gpurun_out/r02d_bench_alias.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02d_bench_alias.json').read().strip().splitlines()[-1])
print('alias', 'ms/step', round(d['ms_per_step'],4), 'svgf', round(d['svgf']['ms_per_frame'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), [ (k['kernel'][:24], round(k['ms']*1e3,1)) for k in d['kernels']])
PY
