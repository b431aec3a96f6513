#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=20 -s > gpurun_out/r01d_pytest.log 2>&1
tail -4 gpurun_out/r01d_pytest.log
grep "^\[parity\]\|^\[raytraced\|^\[gbuffer\|^\[reflection" gpurun_out/r01d_pytest.log | sort | uniq | head -60
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r01d_bench.json 2> gpurun_out/r01d_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r01d_bench.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
for k, v in d["next_rows"].items():
    print(k, round(v["ms"], 4), "ms", round(v["frac"], 4))
for k in d["kernels"]:
    print(k["kernel"], round(k["ms"], 4))
PY
