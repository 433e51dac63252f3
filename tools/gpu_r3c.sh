#!/bin/bash
TAG=${1:-r3c}; O=gpurun_out/$TAG; mkdir -p $O
timeout 200 python tools/time_small.py > $O/small.txt 2>&1; cat $O/small.txt
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "stencil" > $O/pytest_ops.log 2>&1; tail -2 $O/pytest_ops.log
