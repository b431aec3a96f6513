#!/bin/bash
# Round 2, call e: node test with FFMA2 (near/far pairs): parity of every ray test, ray-pass timings, bench.
mkdir -p gpurun_out
python -m pytest tests/test_rt_gpu.py tests/test_golden.py tests/test_textures_gpu.py tests/test_raytraced_path_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q --maxfail=30 > gpurun_out/r02e_pytest.log 2>&1
tail -6 gpurun_out/r02e_pytest.log
VHR_RAYGEN_VARIANT=0 python tools/time_trace.py 3000000 1920 1080 10 > gpurun_out/r02e_trace.log 2>&1; cat gpurun_out/r02e_trace.log
timeout 600 python bench.py --svgf alias --no-cpu-baseline > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02e_bench.json').read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],4), 'svgf', round(d['svgf']['ms_per_frame'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'rt', d['rt_pass'])
PY
