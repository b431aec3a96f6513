#!/bin/bash
# Round 2, call as: bench lines of the other BASELINE configs on one GPU (head commit).
mkdir -p gpurun_out
for wl in shadows_1080p_260ktri ao4_temporal_1080p_1Mtri full_frame_4k_3Mtri; do
timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-strong > gpurun_out/r02as_bench_$wl.json 2> gpurun_out/r02as_bench_$wl.err; python - $wl <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02as_bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), d['unit'], 'svgf', round(d.get('svgf',{}).get('ms_per_frame',0),4), 'e2e', round(d['e2e']['value'],1))
PY
done
