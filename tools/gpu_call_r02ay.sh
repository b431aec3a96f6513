#!/bin/bash
# Round 2, call ay: SSR with conservative step skipping from per-tile depth bounds: parity (bit-exactness) and time, against VHR_SSR_SKIP=0.
mkdir -p gpurun_out
T=gpurun_out/r02ay
python -m pytest tests/test_ssr_gpu.py tests/test_golden.py tests/test_baseline_configs_gpu.py tests/test_host_gpu.py tests/test_partition_gpu.py -m gpu -q -s -k "ssr or golden or next_rows or partition or host" > ${T}_pytest.log 2>&1; tail -2 ${T}_pytest.log; grep "parity\].*ssr" ${T}_pytest.log | cut -c1-200
for v in 1 0 1; do
VHR_SSR_SKIP=$v python bench.py --no-strong --no-cpu-baseline --steps 6 --warmup 3 > ${T}_bench_s$v.json 2> ${T}_bench_s$v.err; python - $v <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02ay_bench_s{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('ssr skip',sys.argv[1],'ssr ms', round(d['next_rows']['ssr']['ms'],3))
PY
done
