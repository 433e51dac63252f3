#!/bin/bash
TAG=${1:-r3e}; O=gpurun_out/$TAG; mkdir -p $O
for L in "128 192 5 2 0 64 64 16 0 0" "128 128 5 2 0 64 64 16 0 0" "128 128 5 2 0 128 128 16 1 0" "128 128 5 2 1 128 128 16 2 0" "128 960 5 1 0 32 32 16 0 1" "128 128 5 2 0 256 256 16 1 0"; do
  timeout 120 python tools/time_layer.py $L 10 >> $O/t.txt 2>&1
done
HESIC_TC_SINGLE_CTA=1 timeout 120 python tools/time_layer.py 128 192 5 2 0 64 64 16 0 0 10 >> $O/t.txt 2>&1
cat $O/t.txt
