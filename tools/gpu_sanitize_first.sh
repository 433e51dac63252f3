#!/bin/bash
TAG=${1:-r3h}; O=gpurun_out/$TAG; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "first_analysis or perspective or max_pool" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
L="3 128 5 2 0 96 80 2 1 0 1"
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/run_layer.py $L > $O/memcheck_first.log 2>&1; tail -4 $O/memcheck_first.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/run_layer.py $L > $O/racecheck_first.log 2>&1; tail -4 $O/racecheck_first.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python tools/run_layer.py $L > $O/synccheck_first.log 2>&1; tail -4 $O/synccheck_first.log
