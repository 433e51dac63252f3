#!/bin/bash
TAG=${1:-r3b}; O=gpurun_out/$TAG; mkdir -p $O
timeout 200 python tools/time_small.py > $O/small.txt 2>&1; cat $O/small.txt
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "stencil or conv" > $O/pytest_ops.log 2>&1; tail -3 $O/pytest_ops.log
timeout 300 python tools/layer_times.py 16 hesic 3 > $O/layers_hesic.txt 2>&1; head -3 $O/layers_hesic.txt; grep "128->192\|6->3" $O/layers_hesic.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --cpu-iters 3 > $O/bench.json 2> $O/bench.err; cut -c1-200 $O/bench.json
