#!/bin/bash
# Round 2, call ak: SSAO with the dividend range tests of the unprojection vouched for by the host.
mkdir -p gpurun_out
T=gpurun_out/r02ak
python -m pytest tests/test_ssao_gpu.py tests/test_golden.py tests/test_baseline_configs_gpu.py tests/test_host_gpu.py tests/test_partition_gpu.py -m gpu -q -s -k "ssao or golden or next_rows or partition or host" > ${T}_pytest.log 2>&1; tail -2 ${T}_pytest.log; grep "parity\].*ssao raw" ${T}_pytest.log | cut -c1-200
for v in 0 0; do
python bench.py --no-strong --no-cpu-baseline --steps 6 --warmup 3 > ${T}_bench.json 2> ${T}_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02ak_bench.json').read().strip().splitlines()[-1])
print('ssao us', round(d['next_rows']['ssao']['ms']*1e3,1), 'blur', round(d['next_rows']['ssao_blur']['ms']*1e3,1), 'ssr ms', round(d['next_rows']['ssr']['ms'],3))
PY
done
