#!/bin/bash
# Round 2, call ab (2 GPUs): head commit on the multi-GPU paths — two-process IPC partition test, partition tests, the N = 2 bench line of both arms.
mkdir -p gpurun_out
T=gpurun_out/r02ab
python -m pytest tests/test_partition_multiprocess_gpu.py tests/test_partition_gpu.py -m gpu -q --maxfail=30 -s > ${T}_pytest.log 2>&1; tail -3 ${T}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > ${T}_bench_n2.json 2> ${T}_bench_n2.err
tail -3 ${T}_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > ${T}_bench_reference_n2.json 2> ${T}_bench_reference_n2.err
python - <<PY
import json
d=json.loads(open('${T}_bench_n2.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1))
e=d['e2e']; print('e2e', round(e['value'],1), 'ms', round(e['ms_per_step'],4), 'camera_in ms', round(e['camera_in']['ms_per_step'],4), round(e['camera_in']['value'],1))
s=d.get('strong_4k') or {}; print({k:s.get(k) for k in ('ms_one_gpu','ms_per_frame','speedup','value','scaling')})
r=json.loads(open('${T}_bench_reference_n2.json').read().strip().splitlines()[-1]); print('reference N=2', r.get('value'), r.get('n_gpus'), r['cpu_baseline']['cores'])
PY
