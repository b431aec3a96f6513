#!/bin/bash
# Round 2, call aa: temporal kernel built for 6 / 8 resident blocks per SM (40 / 32 registers, spills) against the default (48 registers, 5 blocks).
mkdir -p gpurun_out
T=gpurun_out/r02aa
for rep in 1 2; do for k in 1 6 8; do
if [ $k = 1 ]; then unset VHR_LIB_PATH; else export VHR_LIB_PATH=$PWD/build/ab/libvhr_b200_temporal$k.so; fi
python bench.py --no-strong --no-cpu-baseline --steps 20 --warmup 5 > ${T}_bench_$k.json 2> ${T}_bench_$k.err; python - $k <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02aa_bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('temporal min blocks',sys.argv[1],'ms/step',round(d['ms_per_step'],4),'svgf', round(d['svgf']['ms_per_frame'],4))
PY
done; done
