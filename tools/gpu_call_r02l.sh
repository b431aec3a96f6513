#!/bin/bash
# Round 2, call l: two frames in flight re-measured on the round-2 kernels; partition test after the alias-resolution change.
mkdir -p gpurun_out
python -m pytest tests/test_partition_gpu.py tests/test_svgf_gpu.py tests/test_frames_in_flight_gpu.py -m gpu -q > gpurun_out/r02l_pytest.log 2>&1; tail -2 gpurun_out/r02l_pytest.log
for f in 1 2; do
  timeout 600 python bench.py --frames-in-flight $f --no-cpu-baseline --steps 200 > gpurun_out/r02l_bench_fif$f.json 2> gpurun_out/r02l_bench_fif$f.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02l_bench_fif$f.json').read().strip().splitlines()[-1])
print('fif $f: ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],4))
PY
done
