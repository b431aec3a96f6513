#!/bin/bash
# Round 2, call ax (8 GPUs): the N = 8 and N = 4 bench lines of the head commit (weak views + strong 4K fused partition in one line), both arms at N = 8.
mkdir -p gpurun_out
T=gpurun_out/r02ax
for n in 8 4; do
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
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > ${T}_bench_reference_n8.json 2> ${T}_bench_reference_n8.err; cut -c1-200 ${T}_bench_reference_n8.json | tail -1
