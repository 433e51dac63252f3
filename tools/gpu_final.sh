#!/bin/bash
# final check of the committed build: GPU tests, smoke, bench line (default arguments and --steps 20)
TAG=${1:-final}; O=gpurun_out/$TAG; mkdir -p $O
rm -f gpurun_out/parity_stats.jsonl
timeout 1800 python -m pytest tests -m gpu -q --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
cp gpurun_out/parity_stats.jsonl $O/ 2>/dev/null
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_hesic.json 2> $O/bench_hesic.err; echo "bench rc=$?" >> $O/bench_hesic.err
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log; cut -c1-220 $O/bench_hesic.json; cut -c1-220 $O/bench_default.json
exit 0
