#!/bin/bash
# Round 2, call o: SSAO sample loop with the fixed-point filter coordinate, sincos_turn and the continuous part on fast math.
mkdir -p gpurun_out
python -m pytest tests/test_ssao_gpu.py tests/test_golden.py tests/test_host_gpu.py tests/test_partition_gpu.py -m gpu -q --maxfail=30 -s > gpurun_out/r02o_pytest.log 2>&1; tail -3 gpurun_out/r02o_pytest.log
grep "parity\] ssao\|parity\] SSAO\|beyond" gpurun_out/r02o_pytest.log | head -20
python bench.py --no-strong --steps 10 --warmup 3 > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02o_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step')}, d.get('next_rows_ms') or {k:v for k,v in d.items() if 'next' in k})
PY
