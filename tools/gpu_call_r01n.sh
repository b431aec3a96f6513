#!/bin/bash
# the driver's N=2 launch of both arms (views per rank, weak scaling)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 \
  > gpurun_out/r01n_bench_n2.json 2> gpurun_out/r01n_bench_n2.err
echo "rc=$?"; tail -c 600 gpurun_out/r01n_bench_n2.json | cut -c1-600; tail -5 gpurun_out/r01n_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 \
  > gpurun_out/r01n_ref_n2.json 2> gpurun_out/r01n_ref_n2.err
echo "rc=$?"; cut -c1-300 gpurun_out/r01n_ref_n2.json; tail -3 gpurun_out/r01n_ref_n2.err
