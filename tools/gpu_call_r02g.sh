#!/bin/bash
# Round 2, call g (2 GPUs): the N = 2 bench line (weak views + strong 4K partition), the reference arm at N = 2, the tests that changed.
mkdir -p gpurun_out
python -m pytest tests/test_baseline_configs_gpu.py tests/test_rt_gpu.py -m gpu -q --maxfail=30 -s > gpurun_out/r02g_pytest.log 2>&1
tail -4 gpurun_out/r02g_pytest.log; grep -h "\[masks\]\|\[config4\]\|\[config5\]" gpurun_out/r02g_pytest.log | cut -c1-220
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r02g_bench_n2.json 2> gpurun_out/r02g_bench_n2.err
tail -5 gpurun_out/r02g_bench_n2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02g_bench_n2.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1))
e=d['e2e']; print('e2e', round(e['value'],1), 'ms', round(e['ms_per_step'],4), 'serial ms', round(e['serial_ms_per_step'],4), 'camera_in ms', round(e['camera_in']['ms_per_step'],4), round(e['camera_in']['value'],1))
print(json.dumps(d.get('strong_4k'))[:1500])
PY
