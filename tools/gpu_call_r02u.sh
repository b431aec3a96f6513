#!/bin/bash
# Round 2, call u: SSAO default with the shared-reciprocal exact division (no __fdiv_rn slow path for sky samples); no-surface short cut restructured.
mkdir -p gpurun_out
T=gpurun_out/r02u
python -m pytest tests/test_ssao_gpu.py tests/test_golden.py tests/test_host_gpu.py tests/test_partition_gpu.py -m gpu -q --maxfail=30 -s > ${T}_pytest.log 2>&1; tail -2 ${T}_pytest.log
grep "parity\] ssao raw\|beyond" ${T}_pytest.log | grep -v print | head
VHR_SSAO_VARIANT=8 python -m pytest tests/test_ssao_gpu.py -m gpu -q --maxfail=30 -s > ${T}_pytest_v8.log 2>&1; tail -2 ${T}_pytest_v8.log
grep "parity\] ssao raw\|beyond" ${T}_pytest_v8.log | grep -v print | head
for v in 0 4 8 1; do
VHR_SSAO_VARIANT=$v python bench.py --no-strong --no-cpu-baseline --steps 10 --warmup 3 > ${T}_bench_v$v.json 2> ${T}_bench_v$v.err; python - $v <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02u_bench_v{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('variant',sys.argv[1],'ms/step',round(d['ms_per_step'],4), 'ssao us', round(d['next_rows']['ssao']['ms']*1e3,1))
PY
done
python -m pytest tests/test_svgf_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q --maxfail=30 > ${T}_pytest_svgf.log 2>&1; tail -2 ${T}_pytest_svgf.log
for k in 0 1; do
VHR_SVGF_SKIP_NO_SURFACE=$k python bench.py --no-strong --no-cpu-baseline --steps 20 --warmup 5 > ${T}_bench_skip$k.json 2> ${T}_bench_skip$k.err; python - $k <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02u_bench_skip{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('skip',sys.argv[1],'ms/step',round(d['ms_per_step'],4),'svgf', round(d['svgf']['ms_per_frame'],4))
PY
done
