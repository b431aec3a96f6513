#!/bin/bash
# Round 2, call v: SSR with div_exact / sqrt_exact / fixed-point filter coordinate (parity + time); A/B of the SVGF kernels against the
# previous build of svgf_kernels.cu (build/ab/libvhr_b200_oldsvgf.so, same other objects).
mkdir -p gpurun_out
T=gpurun_out/r02v
python -m pytest tests/test_ssr_gpu.py tests/test_composition_gpu.py tests/test_textures_gpu.py tests/test_golden.py -m gpu -q --maxfail=30 -s > ${T}_pytest.log 2>&1; tail -2 ${T}_pytest.log
grep "parity\]" ${T}_pytest.log | grep -i "ssr" | head
for rep in 1 2; do for k in old new; do
if [ $k = old ]; then export VHR_LIB_PATH=$PWD/build/ab/libvhr_b200_oldsvgf.so; else unset VHR_LIB_PATH; fi
python bench.py --no-strong --no-cpu-baseline --steps 20 --warmup 5 > ${T}_bench_$k.json 2> ${T}_bench_$k.err; python - $k <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02v_bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('svgf build',sys.argv[1],'ms/step',round(d['ms_per_step'],4),'svgf', round(d['svgf']['ms_per_frame'],4), 'ssr ms', round(d['next_rows']['ssr']['ms'],3), 'ssao us', round(d['next_rows']['ssao']['ms']*1e3,1))
PY
done; done
unset VHR_LIB_PATH
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"atrous|svgf_" -s 24 -c 36 --csv --log-file ${T}_launches_new.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-strong > ${T}_ncu_new.log 2>&1
VHR_LIB_PATH=$PWD/build/ab/libvhr_b200_oldsvgf.so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"atrous|svgf_" -s 24 -c 36 --csv --log-file ${T}_launches_old.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-strong > ${T}_ncu_old.log 2>&1
python - <<'PY'
import csv,collections
for k in ('old','new'):
    agg=collections.defaultdict(list)
    for r in csv.reader(open(f'gpurun_out/r02v_launches_{k}.csv')):
        if len(r)>14 and r[-2]=='ns': agg[r[4][:48]].append(float(r[-1])/1e3)
    print(k, {n:round(sum(v)/len(v),1) for n,v in agg.items()})
PY
