#!/bin/bash
TAG=${1:-skip}; O=gpurun_out/$TAG; mkdir -p $O
for L in "128 128 5 2 0 256 256 16 0 0" "128 960 5 1 0 32 32 16 0 1"; do
  echo "== $L" >> $O/t.txt
  for v in "HESIC_TC_PAIR_NO_COLS=1" "A=1"; do
    for sk in 0 1 2 3; do
      echo "-- $v skip=$sk" >> $O/t.txt
      env $v HESIC_TC_DBG_SKIP=$sk timeout 120 python tools/time_layer.py $L 2>&1 | tail -1 >> $O/t.txt
    done
  done
done
cat $O/t.txt
