#!/bin/bash
# Round 2, call bc: verification of the head commit — GPU suite, smoke, both bench arms, views64, launch list.
mkdir -p gpurun_out
T=gpurun_out/r02bc
python -m pytest tests -m gpu -q --maxfail=20 > ${T}_pytest.log 2>&1
tail -3 ${T}_pytest.log; grep -h "ssao raw" ${T}_pytest.log | head -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${T}_smoke.log 2>&1; tail -2 ${T}_smoke.log
timeout 900 python bench.py > ${T}_bench.json 2> ${T}_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > ${T}_bench_reference.json 2>> ${T}_bench.err
timeout 600 python bench.py --workload views64_1080p_3Mtri --steps 64 > ${T}_bench_views64.json 2>> ${T}_bench.err; cut -c1-300 ${T}_bench_views64.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"raygen_kernel|atrous|svgf_|composition_kernel|ssao|depth_quads|ssr_kernel|gbuffer_kernel" -s 24 -c 120 --csv --log-file ${T}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > ${T}_ncu_launches.log 2>&1
python - <<PY
import json
d=json.loads(open('${T}_bench.json').read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'svgf', round(d['svgf']['ms_per_frame'],4), 'e2e', round(d['e2e']['value'],1), 'launches', d.get('gpu_launches'))
print({k:round(v['ms']*1e3,1) for k,v in d.get('next_rows',{}).items()})
r=json.loads(open('${T}_bench_reference.json').read().strip().splitlines()[-1]); print('reference', r['value'], r['cpu_baseline']['cores'], r['cpu_baseline'].get('svgf_ms_per_full_frame_measured'))
PY
