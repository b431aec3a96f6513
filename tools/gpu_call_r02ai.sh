#!/bin/bash
# Round 2, call ai: tiled SSAO blur (64 x 32 tile, register blocking) against the first kernel.
mkdir -p gpurun_out
T=gpurun_out/r02ai
python -m pytest tests/test_ssao_gpu.py tests/test_golden.py tests/test_baseline_configs_gpu.py tests/test_host_gpu.py tests/test_partition_gpu.py -m gpu -q -s -k "ssao or golden or next_rows or partition or host" > ${T}_pytest.log 2>&1; tail -2 ${T}_pytest.log; grep "parity\].*blur" ${T}_pytest.log | cut -c1-200
for v in 0 1 0; do
VHR_SSAO_BLUR_VARIANT=$v python bench.py --no-strong --no-cpu-baseline --steps 6 --warmup 3 > ${T}_bench_v$v.json 2> ${T}_bench_v$v.err; python - $v <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02ai_bench_v{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('blur variant',sys.argv[1],'blur us', round(d['next_rows']['ssao_blur']['ms']*1e3,2))
PY
done
