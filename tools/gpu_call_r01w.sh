#!/bin/bash
# launch list of the bench frames on the last commit of the round
mkdir -p gpurun_out
T=gpurun_out/r01w
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"raygen_kernel|atrous|svgf_temporal|composition_kernel|ssao|ssr_kernel|gbuffer_kernel" -s 24 -c 120 --csv --log-file ${T}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > ${T}_ncu_launches.log 2>&1
tail -c 300 ${T}_ncu_launches.log
