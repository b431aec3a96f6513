#!/bin/bash
# Round 2, call r: SSAO exact default (grouped quad loads); variants 4 = sincos_turn, 8 = fast continuous part, 12, 2 = no quads, 1 = old loop.
mkdir -p gpurun_out
python -m pytest tests/test_ssao_gpu.py tests/test_golden.py tests/test_host_gpu.py tests/test_partition_gpu.py tests/test_svgf_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q --maxfail=30 -s > gpurun_out/r02r_pytest.log 2>&1; tail -3 gpurun_out/r02r_pytest.log
grep "parity\] ssao\|beyond" gpurun_out/r02r_pytest.log | grep -v print | head -20
for v in 0 4 8 12 2 1; do
VHR_SSAO_VARIANT=$v python bench.py --no-strong --steps 10 --warmup 3 > gpurun_out/r02r_bench_v$v.json 2> gpurun_out/r02r_bench_v$v.err; python - $v <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02r_bench_v{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('variant',sys.argv[1],{k:d.get(k) for k in ('value','ms_per_step')}, 'ssao ms', d['next_rows']['ssao']['ms'], 'svgf', d.get('svgf',{}).get('ms_per_frame'), {k:v for k,v in d.get('passes_ms',{}).items()} if 'passes_ms' in d else '')
PY
done
