#!/bin/bash
# Round 2, call ah: SSR lane refill again, now that a probe's tap is one load (quad image).
mkdir -p gpurun_out
T=gpurun_out/r02ah
for r in 22 42 82; do
VHR_SSR_VARIANT=1 VHR_SSR_REGION=$r python bench.py --no-strong --no-cpu-baseline --steps 6 --warmup 3 > ${T}_bench_r$r.json 2> ${T}_bench_r$r.err; python - $r <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02ah_bench_r{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('ssr refill region',sys.argv[1],'ssr ms', round(d['next_rows']['ssr']['ms'],3))
PY
done
VHR_BENCH_ROW_REPS=1 timeout 900 ncu --set full --clock-control none -k regex:"ssr_kernel" -s 1 -c 1 -o /tmp/r02ah_ssr python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong > ${T}_ncu_ssr.log 2>&1
ncu -i /tmp/r02ah_ssr.ncu-rep --page raw --csv > ${T}_ncu_ssr_raw.csv 2>> ${T}_ncu_ssr.log
