#!/bin/bash
mkdir -p gpurun_out/j17
O=gpurun_out/j17
HESIC_ONE_STREAM=1 timeout 300 python tools/kernel_breakdown.py hesic 16 > $O/hesic_breakdown.txt 2>&1
HESIC_ONE_STREAM=1 timeout 300 python tools/layer_times.py 16 hesic 3 > $O/layer_times.txt 2>&1
head -2 $O/hesic_breakdown.txt; cat $O/layer_times.txt
