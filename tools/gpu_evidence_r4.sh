#!/bin/bash
# r04 (third session of round 2) evidence in ONE gpurun call: tests, smoke, bench (both arms), launch list, stream A/B.
#   tools/gpu_evidence_r4.sh <tag>
TAG=${1:-ev4}; O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt 2>&1
rm -f gpurun_out/parity_stats.jsonl
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
cp gpurun_out/parity_stats.jsonl $O/ 2>/dev/null
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_hesic.json 2> $O/bench_hesic.err; echo "bench rc=$?" >> $O/bench_hesic.err
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --cpu-iters 1 --no-extras --sustain-s 0.01 > $O/bench_ncu.log 2>&1
timeout 120 python tools/time_forward.py hesic 16 5 > $O/time_forward.txt 2>&1
timeout 120 python tools/time_forward.py hesic_plus 16 5 >> $O/time_forward.txt 2>&1
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log; cut -c1-300 $O/bench_hesic.json; tail -2 $O/bench_hesic.err; cut -c1-200 $O/bench_default.json; cut -c1-200 $O/bench_reference.json; cat $O/time_forward.txt
exit 0
