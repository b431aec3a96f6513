#!/bin/bash
# Round-1 final capture: GPU suite, smoke(), both bench arms, ncu launch list of the bench command, one ncu --set full capture of a
# frame (raygen + temporal + 5 a-trous) and of the next-row kernels. Only CSV extracts come back (the .ncu-rep files are large).
mkdir -p gpurun_out
T=gpurun_out/r01p
python -m pytest tests -m gpu -q --maxfail=20 > ${T}_pytest.log 2>&1
tail -3 ${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${T}_smoke.log 2>&1; tail -2 ${T}_smoke.log
timeout 900 python bench.py > ${T}_bench.json 2> ${T}_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > ${T}_bench_reference.json 2>> ${T}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file ${T}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > ${T}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"raygen_kernel|atrous_pair|svgf_temporal" -s 31 -c 7 -o /tmp/r01p_frame \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > ${T}_ncu_frame.log 2>&1
ncu -i /tmp/r01p_frame.ncu-rep --page raw --csv > ${T}_ncu_frame_raw.csv 2>> ${T}_ncu_frame.log
VHR_BENCH_ROW_REPS=1 timeout 900 ncu --set full --clock-control none -k regex:"composition_kernel|ssao_kernel|ssao_blur_kernel|gbuffer_kernel" -s 3 -c 12 -o /tmp/r01p_rows \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > ${T}_ncu_rows.log 2>&1
ncu -i /tmp/r01p_rows.ncu-rep --page raw --csv > ${T}_ncu_rows_raw.csv 2>> ${T}_ncu_rows.log
ls -la gpurun_out/ | grep r01p; ls -la /tmp/*.ncu-rep
