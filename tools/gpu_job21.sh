#!/bin/bash
mkdir -p gpurun_out/j21
O=gpurun_out/j21
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
HESIC_ONE_STREAM=1 timeout 300 python tools/layer_times.py 16 hesic 3 > $O/layer_times.txt 2>&1; head -1 $O/layer_times.txt; grep -n "warp\|spatial\|convert\|gaussian" $O/layer_times.txt
timeout 300 python tools/layer_times.py 16 hesic 3 2>&1 | head -1
