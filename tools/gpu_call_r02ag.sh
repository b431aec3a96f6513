#!/bin/bash
# Round 2, call ag: SSR depth taps from the quad image.
mkdir -p gpurun_out
T=gpurun_out/r02ag
python -m pytest tests/test_ssr_gpu.py tests/test_golden.py tests/test_baseline_configs_gpu.py tests/test_host_gpu.py tests/test_partition_gpu.py tests/test_ssao_gpu.py -m gpu -q -s -k "ssr or golden or next_rows or partition or host or ssao" > ${T}_pytest.log 2>&1; tail -2 ${T}_pytest.log; grep "parity\].*ssr" ${T}_pytest.log | cut -c1-200
python bench.py --no-strong --no-cpu-baseline --steps 10 --warmup 3 > ${T}_bench.json 2> ${T}_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02ag_bench.json').read().strip().splitlines()[-1])
print('ms/step',round(d['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['next_rows'].items()})
PY
