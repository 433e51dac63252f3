#!/bin/bash
# GPU test suite (+ golden codec files on request):  tools/gpu_tests.sh <tag> [golden] [pytest args...]
TAG=${1:-t}; shift
O=gpurun_out/$TAG
mkdir -p $O
rm -f gpurun_out/parity_stats.jsonl
if [ "$1" == "golden" ]; then
  shift
  timeout 600 python tests/golden/make_codec_golden.py gpurun_out/codec_golden > $O/codec_golden.log 2>&1; echo "golden rc=$?" >> $O/codec_golden.log
  cp gpurun_out/codec_golden/codec_* tests/golden/ 2>/dev/null
fi
timeout 1800 python -m pytest tests -m gpu -q --durations=10 "$@" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
cp gpurun_out/parity_stats.jsonl $O/ 2>/dev/null
tail -5 $O/codec_golden.log 2>/dev/null; grep -E "^(FAILED|ERROR)|passed|failed" $O/pytest_gpu.log | tail -30
