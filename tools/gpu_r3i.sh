#!/bin/bash
TAG=${1:-r3i}; O=gpurun_out/$TAG; mkdir -p $O
timeout 300 python tools/tc_check.py head > $O/check.txt 2>&1; cat $O/check.txt | tail -5
for L in "128 3 5 2 1 256 256 16 2 0" "128 3 5 2 1 256 256 16 0 0"; do timeout 120 python tools/time_layer.py $L 20 >> $O/t.txt 2>&1; done
cat $O/t.txt
