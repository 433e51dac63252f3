#!/bin/bash
# quick check of a change: GPU tests + one bench line:  tools/gpu_check.sh <tag> [bench args...]
TAG=${1:-chk}; shift
O=gpurun_out/$TAG; mkdir -p $O
rm -f gpurun_out/parity_stats.jsonl
timeout 1200 python -m pytest tests -m gpu -q -x --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
cp gpurun_out/parity_stats.jsonl $O/ 2>/dev/null
timeout 600 python bench.py --steps 20 --warmup 5 "$@" > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
grep -E "^(FAILED|ERROR)|passed|failed|Error" $O/pytest_gpu.log | tail -15; cut -c1-300 $O/bench.json; tail -3 $O/bench.err
exit 0
