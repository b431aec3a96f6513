#!/bin/bash
# Round 2, call ae: SSR with lane refill (default) against the per-pixel kernel (VHR_SSR_VARIANT=1): parity and time.
mkdir -p gpurun_out
T=gpurun_out/r02ae
for v in 0 1; do
VHR_SSR_VARIANT=$v python -m pytest tests/test_ssr_gpu.py tests/test_golden.py tests/test_baseline_configs_gpu.py tests/test_host_gpu.py tests/test_partition_gpu.py -m gpu -q -s -k "ssr or golden or next_rows or partition or host" > ${T}_pytest_v$v.log 2>&1; tail -2 ${T}_pytest_v$v.log; grep "parity\].*ssr" ${T}_pytest_v$v.log | cut -c1-200
VHR_SSR_VARIANT=$v python bench.py --no-strong --no-cpu-baseline --steps 10 --warmup 3 > ${T}_bench_v$v.json 2> ${T}_bench_v$v.err; python - $v <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02ae_bench_v{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('ssr variant',sys.argv[1],'ms/step',round(d['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['next_rows'].items()})
PY
done
