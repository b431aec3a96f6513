#!/bin/bash
# Round-1 capture (second attempt: the .ncu-rep files stay on the box, only CSV extracts come back).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r01f_pytest.log 2>&1
tail -3 gpurun_out/r01f_pytest.log
timeout 900 python bench.py > gpurun_out/r01f_bench.json 2> gpurun_out/r01f_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01f_bench_reference.json 2>> gpurun_out/r01f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/r01f_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01f_ncu_launches.log 2>&1
# one frame of the step: raygen + temporal + 5 a-trous (frame 5 of 6)
timeout 900 ncu --set full --clock-control none -k regex:"raygen_kernel|atrous_pair|svgf_temporal" -s 31 -c 7 -o /tmp/r01f_frame \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01f_ncu_frame.log 2>&1
ncu -i /tmp/r01f_frame.ncu-rep --page raw --csv > gpurun_out/r01f_ncu_frame_raw.csv 2>> gpurun_out/r01f_ncu_frame.log
# the next-row kernels, 4 launches each
VHR_BENCH_ROW_REPS=1 timeout 900 ncu --set full --clock-control none -k regex:"composition_kernel|ssao_kernel|ssao_blur_kernel|ssr_kernel|gbuffer_kernel" -s 3 -c 20 -o /tmp/r01f_rows \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01f_ncu_rows.log 2>&1
ncu -i /tmp/r01f_rows.ncu-rep --page raw --csv > gpurun_out/r01f_ncu_rows_raw.csv 2>> gpurun_out/r01f_ncu_rows.log
ls -la gpurun_out/ /tmp/*.ncu-rep
