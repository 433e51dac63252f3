#!/bin/bash
# Round-2 GPU evidence, one gpurun call: tests, smoke, bench lines (both arms), per-layer table, launch list.
#   tools/gpu_r2.sh <tag> [quick]
TAG=${1:-r2}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt 2>&1
rm -f gpurun_out/parity_stats.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
cp gpurun_out/parity_stats.jsonl $O/ 2>/dev/null
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_hesic.json 2> $O/bench_hesic.err; echo "bench rc=$?" >> $O/bench_hesic.err
if [ "$2" != "quick" ]; then
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
  HESIC_ONE_STREAM=1 timeout 300 python tools/layer_times.py 16 hesic 3 > $O/layer_times_hesic.txt 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --cpu-iters 1 --no-extras --sustain-s 0.01 > $O/bench_ncu.log 2>&1
fi
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log; cut -c1-600 $O/bench_hesic.json; tail -3 $O/bench_hesic.err
