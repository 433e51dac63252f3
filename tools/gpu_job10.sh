#!/bin/bash
mkdir -p gpurun_out/j10
O=gpurun_out/j10
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python tools/layer_times.py 16 hesic 3 > $O/layer_wide.txt 2>&1
HESIC_TC_NARROW=1 timeout 300 python tools/layer_times.py 16 hesic 3 > $O/layer_narrow.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-iters 2 > $O/bench.json 2> $O/bench.err
tail -3 $O/pytest_gpu.log; head -12 $O/layer_wide.txt; head -8 $O/layer_narrow.txt; cut -c1-300 $O/bench.json
