#!/bin/bash
# two frames in flight: parity tests + bench with 1 and 2 frames in flight on the same box
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_frames_in_flight_gpu.py tests/test_transfer_gpu.py -m gpu -x -q > gpurun_out/r01o_pytest.log 2>&1
tail -15 gpurun_out/r01o_pytest.log
for f in 1 2; do
  timeout 600 python bench.py --frames-in-flight $f --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r01o_bench_fif$f.json 2> gpurun_out/r01o_bench_fif$f.err
  tail -3 gpurun_out/r01o_bench_fif$f.err
  python - <<P
import json
d=[json.loads(l) for l in open("gpurun_out/r01o_bench_fif$f.json") if l.startswith("{")][-1]
print("fif $f", {k:round(d[k],4) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],4), "serial e2e", round(d["e2e"]["serial_ms_per_step"],4), "rt", round(d["rt_pass"]["ms"],4), "svgf", round(d["svgf"]["ms_per_frame"],4), "checksums", d["e2e"]["checksum"], d["e2e"]["serial_checksum"])
P
done
