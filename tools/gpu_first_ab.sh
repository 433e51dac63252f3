#!/bin/bash
O=gpurun_out/$1; mkdir -p $O
L="3 128 5 2 0 512 512 16 1 0"
for v in elect lane0 elect lane0; do
  cp hesic_b200/lib/lib_$v.so hesic_b200/lib/libhesic_b200.so
  echo "== $v" >> $O/t.txt
  timeout 120 python tools/time_layer.py $L 20 >> $O/t.txt 2>&1
  HESIC_TC_FIRST_DBG=1 timeout 120 python tools/time_layer.py $L 20 >> $O/t.txt 2>&1
done
cat $O/t.txt
