#!/bin/bash
mkdir -p gpurun_out/j12
O=gpurun_out/j12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --cpu-iters 2 > $O/bench_n2.json 2> $O/bench_n2.err
echo "rc=$?"; cut -c1-500 $O/bench_n2.json; tail -5 $O/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
echo "rc=$?"; cut -c1-300 $O/bench_ref_n2.json
