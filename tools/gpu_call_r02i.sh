#!/bin/bash
# Round 2, call i (8 GPUs): the N = 8, 4, 2 bench lines (weak views + strong 4K fused partition in one line) and the two-process partition test.
mkdir -p gpurun_out
T=gpurun_out/r02i
nvidia-smi topo -m > ${T}_topo.txt 2>&1
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 50 --warmup 5 > ${T}_bench_n$n.json 2> ${T}_bench_n$n.err
  tail -2 ${T}_bench_n$n.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open('${T}_bench_n$n.json').read().strip().splitlines()[-1])
    e=d['e2e']; s=d.get('strong_4k') or {}
    print('N', d['n_gpus'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), '| e2e', round(e['value'],1), 'ms', round(e['ms_per_step'],4), '| camera_in', round(e['camera_in']['value'],1), 'ms', round(e['camera_in']['ms_per_step'],4))
    print('  strong_4k: 1gpu', round(s.get('ms_per_frame_1gpu',0),3), 'N', round(s.get('ms_per_frame',0),3), 'speedup', round(s.get('speedup_vs_1gpu',0),3), 'e2e ms', round(s.get('e2e',{}).get('ms_per_frame',0),3), 'min-rank ms', round(s.get('ms_per_frame_min_over_ranks',0),3))
except Exception as ex:
    print('N $n failed', ex)
PY
done
timeout 600 python -m pytest tests/test_partition_multiprocess_gpu.py -m gpu -q -s 2>&1 | tail -8 | cut -c1-300 | tee ${T}_pytest_2proc.log
