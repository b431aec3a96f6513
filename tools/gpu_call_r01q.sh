#!/bin/bash
# bilinear_setup in-range fast path: tests of every kernel that samples through it + bench rows; launch list of the bench frames only
mkdir -p gpurun_out
T=gpurun_out/r01q
timeout 600 python -m pytest tests/test_ssao_gpu.py tests/test_ssr_gpu.py tests/test_composition_gpu.py tests/test_textures_gpu.py tests/test_golden.py tests/test_partition_gpu.py tests/test_host_gpu.py -m gpu -q --maxfail=20 > ${T}_pytest.log 2>&1
tail -3 ${T}_pytest.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > ${T}_bench.json 2> ${T}_bench.err
python - <<'P'
import json
d=[json.loads(l) for l in open("gpurun_out/r01q_bench.json") if l.startswith("{")][-1]
print({k:round(d[k],4) for k in ("value","ms_per_step")}, {k:round(v["ms"],4) for k,v in d["next_rows"].items()})
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"raygen_kernel|atrous|svgf_temporal|composition_kernel|ssao|ssr_kernel|gbuffer_kernel" -s 24 -c 120 --csv --log-file ${T}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > ${T}_ncu_launches.log 2>&1
tail -2 ${T}_ncu_launches.log | cut -c1-300
