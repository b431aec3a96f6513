#!/bin/bash
# Round 2, call f: the whole GPU suite on the current build + the restructured bench at N = 1.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r02f_pytest.log 2>&1
tail -6 gpurun_out/r02f_pytest.log
timeout 900 python bench.py --steps 100 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -5 gpurun_out/r02f_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02f_bench.json').read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'svgf', round(d['svgf']['ms_per_frame'],4), 'frac_min', round(d['svgf']['frac_of_peak_vs_fused_minimum'],4))
e=d['e2e']; print('e2e', round(e['value'],1), 'ms', round(e['ms_per_step'],4), 'serial ms', round(e['serial_ms_per_step'],4), 'camera_in ms', round(e['camera_in']['ms_per_step'],4), round(e['camera_in']['value'],1))
print('cpu', d.get('cpu_baseline',{}).get('value'), 'roofline', {k:v for k,v in d['roofline'].items() if k in ('kernel','frac','issue')})
print({k:round(v['ms']*1e3,1) for k,v in d.get('next_rows',{}).items()})
PY
