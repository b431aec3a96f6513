#!/bin/bash
# Round 2, call ba: SSR step skipping with the cheaper test and a warp-uniform cooldown (0 / 1 / 3 / 7 rounds), against no skipping.
mkdir -p gpurun_out
T=gpurun_out/r02ba
python -m pytest tests/test_ssr_gpu.py tests/test_golden.py tests/test_baseline_configs_gpu.py tests/test_host_gpu.py -m gpu -q -s -k "ssr or golden or next_rows or host" > ${T}_pytest.log 2>&1; tail -2 ${T}_pytest.log; grep "parity\].*ssr" ${T}_pytest.log | cut -c1-200
for c in noskip 0 1 3 7; do
unset VHR_LIB_PATH; unset VHR_SSR_SKIP
if [ $c = noskip ]; then export VHR_SSR_SKIP=0; elif [ $c != 3 ]; then export VHR_LIB_PATH=$PWD/build/ab/libvhr_b200_ssrc$c.so; fi
python bench.py --no-strong --no-cpu-baseline --steps 6 --warmup 3 > ${T}_bench_$c.json 2> ${T}_bench_$c.err; python - $c <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02ba_bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('ssr cooldown',sys.argv[1],'ssr ms', round(d['next_rows']['ssr']['ms'],3))
PY
done
