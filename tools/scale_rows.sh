#!/bin/bash
# usage: tools/scale_rows.sh "<N list>" [extra bench args]   (run under gpurun --gpus max(N))
# Strong scaling of ONE 4K frame (BASELINE config 4) over row partitions; one JSON line per N into gpurun_out/rows_<tag>_nN.json
NS="$1"; shift
TAG="${TAG:-fused}"
mkdir -p gpurun_out
for N in $NS; do
  if [ "$N" = "1" ]; then
    python bench.py --partition rows --workload full_frame_4k_3Mtri --steps 50 --warmup 5 "$@" > gpurun_out/rows_${TAG}_n1.json 2> gpurun_out/rows_${TAG}_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N \
      --partition rows --workload full_frame_4k_3Mtri --steps 50 --warmup 5 "$@" > gpurun_out/rows_${TAG}_n$N.json 2> gpurun_out/rows_${TAG}_n$N.err
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/rows_${TAG}_n$N.json").read().strip().splitlines()[-1])
    print("${TAG}", $N, "GPUs:", round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 3), "ms/frame", "checksum", d["denoised_checksum"])
except Exception as e:
    print("${TAG}", $N, "FAILED", e); print(open("gpurun_out/rows_${TAG}_n$N.err").read()[-1500:])
PY
done
