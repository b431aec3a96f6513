#!/bin/bash
# Round 2, call x: SSAO default with one fall-back branch per sample, one-period REPEAT in wrap_repeat (SSR), SSR K = 1.
mkdir -p gpurun_out
T=gpurun_out/r02x
python -m pytest tests/test_ssao_gpu.py tests/test_ssr_gpu.py tests/test_golden.py tests/test_host_gpu.py tests/test_partition_gpu.py tests/test_composition_gpu.py tests/test_textures_gpu.py -m gpu -q --maxfail=30 -s > ${T}_pytest.log 2>&1; tail -2 ${T}_pytest.log
grep "parity\] ssao raw\|beyond" ${T}_pytest.log | grep -v print | head
for v in 0 4 0; do
VHR_SSAO_VARIANT=$v python bench.py --no-strong --no-cpu-baseline --steps 10 --warmup 3 > ${T}_bench_v$v.json 2> ${T}_bench_v$v.err; python - $v <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02x_bench_v{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('variant',sys.argv[1],'ms/step',round(d['ms_per_step'],4),'svgf', round(d['svgf']['ms_per_frame'],4), 'ssr ms', round(d['next_rows']['ssr']['ms'],3), 'ssao us', round(d['next_rows']['ssao']['ms']*1e3,1))
PY
done
