#!/bin/bash
mkdir -p gpurun_out/n4
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 20 --warmup 5 --cpu-iters 2 > gpurun_out/n4/bench_n4.json 2> gpurun_out/n4/bench_n4.err
echo "rc=$?"; cut -c1-400 gpurun_out/n4/bench_n4.json; tail -3 gpurun_out/n4/bench_n4.err
