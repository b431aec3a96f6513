#!/bin/bash
# ncu re-capture on the last build of the round: launch list of the bench frames + one ncu --set full frame
mkdir -p gpurun_out
T=gpurun_out/r01t
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"raygen_kernel|atrous|svgf_temporal|composition_kernel|ssao|ssr_kernel|gbuffer_kernel" -s 24 -c 120 --csv --log-file ${T}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > ${T}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"raygen_kernel|atrous_pair|svgf_temporal" -s 31 -c 7 -o /tmp/r01t_frame \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > ${T}_ncu_frame.log 2>&1
ncu -i /tmp/r01t_frame.ncu-rep --page raw --csv > ${T}_ncu_frame_raw.csv 2>> ${T}_ncu_frame.log
ls -la gpurun_out/ | grep r01t
