#!/bin/bash
TAG=${1:-pf}; O=gpurun_out/$TAG; mkdir -p $O
for L in "128 128 5 2 0 256 256 16 1 0" "128 128 5 2 1 128 128 16 2 0" "128 960 5 1 0 32 32 16 0 1" "128 128 5 2 0 128 128 16 1 0"; do
  echo "== $L" >> $O/t.txt
  for pf in 0 32 4 1; do
    echo "-- pf=$pf" >> $O/t.txt
    HESIC_TC_PF=$pf timeout 120 python tools/time_layer.py $L 2>&1 | tail -1 >> $O/t.txt
  done
done
cat $O/t.txt
