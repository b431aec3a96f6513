#!/bin/bash
# Round 2, call af: SSR lane refill — region size / gather threshold sweep.
mkdir -p gpurun_out
T=gpurun_out/r02af
for r in 21 22 41 42 44 81 82; do
VHR_SSR_REGION=$r python bench.py --no-strong --no-cpu-baseline --steps 6 --warmup 3 > ${T}_bench_r$r.json 2> ${T}_bench_r$r.err; python - $r <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02af_bench_r{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('ssr region',sys.argv[1],'ssr ms', round(d['next_rows']['ssr']['ms'],3))
PY
done
