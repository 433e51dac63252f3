#!/bin/bash
mkdir -p gpurun_out/n2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 5 --cpu-iters 2 > gpurun_out/n2/bench_n2.json 2> gpurun_out/n2/bench_n2.err
echo "rc=$?"; cut -c1-260 gpurun_out/n2/bench_n2.json
