#!/bin/bash
# Round 2, call s: A/B of the no-surface short cut in the SVGF kernels on the bench frame; ncu of the SSAO loop variants.
mkdir -p gpurun_out
T=gpurun_out/r02s
for rep in 1 2; do for k in 0 1; do
VHR_SVGF_SKIP_NO_SURFACE=$k python bench.py --no-strong --no-cpu-baseline --steps 20 --warmup 5 > ${T}_bench_skip$k.json 2> ${T}_bench_skip$k.err; python - $k <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02s_bench_skip{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('skip',sys.argv[1],'ms/step',round(d['ms_per_step'],4),'svgf', round(d['svgf']['ms_per_frame'],4), {k:v for k,v in d.get('svgf',{}).items() if 'kern' in k or 'iter' in k})
PY
done; done
for v in 1 8 0; do
VHR_SSAO_VARIANT=$v VHR_BENCH_ROW_REPS=1 timeout 600 ncu --set full --clock-control none -k regex:"ssao_kernel" -s 1 -c 2 -o /tmp/r02s_ssao_v$v python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong > ${T}_ncu_ssao_v$v.log 2>&1
ncu -i /tmp/r02s_ssao_v$v.ncu-rep --page raw --csv > ${T}_ncu_ssao_v${v}_raw.csv 2>> ${T}_ncu_ssao_v$v.log
done
ls -la gpurun_out | tail -5
