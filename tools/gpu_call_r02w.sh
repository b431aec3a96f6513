#!/bin/bash
# Round 2, call w: SSR march K probes per round (K = 1 / 2 / 4: build/ab libraries); ncu of ssr_kernel and of the SSAO default.
mkdir -p gpurun_out
T=gpurun_out/r02w
python -m pytest tests/test_ssr_gpu.py tests/test_svgf_gpu.py -m gpu -q --maxfail=30 -s > ${T}_pytest.log 2>&1; tail -2 ${T}_pytest.log
for k in 1 2 4; do
if [ $k = 2 ]; then unset VHR_LIB_PATH; else export VHR_LIB_PATH=$PWD/build/ab/libvhr_b200_ssrk$k.so; fi
python bench.py --no-strong --no-cpu-baseline --steps 10 --warmup 3 > ${T}_bench_k$k.json 2> ${T}_bench_k$k.err; python - $k <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02w_bench_k{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('ssr K',sys.argv[1],'ms/step',round(d['ms_per_step'],4),'svgf', round(d['svgf']['ms_per_frame'],4), 'ssr ms', round(d['next_rows']['ssr']['ms'],3), 'ssao us', round(d['next_rows']['ssao']['ms']*1e3,1))
PY
done
unset VHR_LIB_PATH
VHR_BENCH_ROW_REPS=1 timeout 900 ncu --set full --clock-control none -k regex:"ssao_kernel|ssr_kernel" -s 1 -c 4 -o /tmp/r02w_rows python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong > ${T}_ncu_rows.log 2>&1
ncu -i /tmp/r02w_rows.ncu-rep --page raw --csv > ${T}_ncu_rows_raw.csv 2>> ${T}_ncu_rows.log
