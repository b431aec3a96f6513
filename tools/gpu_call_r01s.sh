#!/bin/bash
# final check of the committed state: GPU suite, smoke(), both bench arms
mkdir -p gpurun_out
T=gpurun_out/r01s
python -m pytest tests -m gpu -q --maxfail=20 > ${T}_pytest.log 2>&1
tail -4 ${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${T}_smoke.log 2>&1; tail -1 ${T}_smoke.log
timeout 900 python bench.py > ${T}_bench.json 2> ${T}_bench.err; tail -2 ${T}_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > ${T}_bench_reference.json 2>> ${T}_bench.err
python - <<'P'
import json
for f in ("gpurun_out/r01s_bench.json","gpurun_out/r01s_bench_reference.json"):
    d=[json.loads(l) for l in open(f) if l.startswith("{")][-1]
    print(f, round(d["value"],2), round(d["ms_per_step"],4), d["e2e"]["value"], d.get("cpu_baseline",{}).get("cores"), d["config"].get("frames_in_flight"))
P
