#!/bin/bash
TAG=${1:-r3f}; O=gpurun_out/$TAG; mkdir -p $O
rm -f gpurun_out/parity_stats.jsonl
timeout 1800 python -m pytest tests -m gpu -q --durations=10 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
cp gpurun_out/parity_stats.jsonl $O/ 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_hesic.json 2> $O/bench_hesic.err; echo "bench rc=$?" >> $O/bench_hesic.err
tail -4 $O/pytest_gpu.log; cut -c1-300 $O/bench_hesic.json; tail -2 $O/bench_hesic.err
