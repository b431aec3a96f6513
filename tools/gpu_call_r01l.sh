#!/bin/bash
# PLOC + axis-sorted slots as defaults: full GPU suite + bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r01l_pytest.log 2>&1
tail -3 gpurun_out/r01l_pytest.log
timeout 900 python bench.py > gpurun_out/r01l_bench.json 2> gpurun_out/r01l_bench.err
python - <<'P'
import json
d=[json.loads(l) for l in open("gpurun_out/r01l_bench.json") if l.startswith("{")][-1]
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["rt_pass"], d["svgf"]["ms_per_frame"], d["bvh"])
print({k:round(v["ms"],4) for k,v in d["next_rows"].items()})
P
