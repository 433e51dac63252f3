#!/bin/bash
mkdir -p gpurun_out/n8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 --cpu-iters 2 > gpurun_out/n8/bench_n8.json 2> gpurun_out/n8/bench_n8.err
echo "rc=$?"; cut -c1-400 gpurun_out/n8/bench_n8.json; tail -3 gpurun_out/n8/bench_n8.err
