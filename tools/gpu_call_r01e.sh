#!/bin/bash
# Final round-1 capture: the whole -m gpu suite, the default bench line, the ncu launch list of the same command and one
# --set full capture of the kernels of a frame + the next-row kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r01e_pytest.log 2>&1
tail -3 gpurun_out/r01e_pytest.log
timeout 900 python bench.py > gpurun_out/r01e_bench.json 2> gpurun_out/r01e_bench.err
tail -c 600 gpurun_out/r01e_bench.json; echo
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01e_bench_reference.json 2>> gpurun_out/r01e_bench.err
tail -c 400 gpurun_out/r01e_bench_reference.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/r01e_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01e_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"raygen_kernel|atrous_pair|svgf_temporal|composition_kernel|ssao_kernel|ssao_blur|ssr_kernel|gbuffer_kernel" -s 24 -c 40 \
    -o gpurun_out/r01e_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01e_ncu_full.log 2>&1
ls -la gpurun_out/r01e_full.ncu-rep
