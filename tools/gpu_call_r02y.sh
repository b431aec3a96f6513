#!/bin/bash
# Round 2, call y: per-triangle any-hit flag (closest-hit traversals skip the material look-up for opaque triangles).
mkdir -p gpurun_out
T=gpurun_out/r02y
python -m pytest tests/test_rt_gpu.py tests/test_raytraced_path_gpu.py tests/test_textures_gpu.py tests/test_golden.py tests/test_host_gpu.py -m gpu -q --maxfail=30 > ${T}_pytest.log 2>&1; tail -2 ${T}_pytest.log
python bench.py --no-strong --no-cpu-baseline --steps 10 --warmup 3 > ${T}_bench.json 2> ${T}_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02y_bench.json').read().strip().splitlines()[-1])
print('ms/step',round(d['ms_per_step'],4),'svgf', round(d['svgf']['ms_per_frame'],4), {k:round(v['ms']*1e3,1) for k,v in d['next_rows'].items()}, 'camera_in', d['e2e'].get('camera_in'))
PY
python bench.py --workload views64_1080p_3Mtri --no-cpu-baseline > ${T}_bench_views64.json 2> ${T}_bench_views64.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02y_bench_views64.json').read().strip().splitlines()[-1])
print('views64', d['value'], d['unit'], d.get('ms_per_step'), {k:d[k] for k in d if 'frames' in k})
PY
