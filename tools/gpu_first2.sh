#!/bin/bash
TAG=${1:-first2}; O=gpurun_out/$TAG; mkdir -p $O
L="3 128 5 2 0 512 512 16 1 0"
for st in 5 4 3 2; do echo "RC16 stages=$st" >> $O/t.txt; HESIC_TC_FIRST_STAGES=$st timeout 120 python tools/time_layer.py $L 10 >> $O/t.txt 2>&1; done
echo "RC16 nostore" >> $O/t.txt; HESIC_TC_FIRST_DBG=1 timeout 120 python tools/time_layer.py $L 10 >> $O/t.txt 2>&1
echo "RC32 nostore" >> $O/t.txt; HESIC_TC_FIRST_RC=32 HESIC_TC_FIRST_DBG=1 timeout 120 python tools/time_layer.py $L 10 >> $O/t.txt 2>&1
echo "RC32 stages=2" >> $O/t.txt; HESIC_TC_FIRST_RC=32 HESIC_TC_FIRST_STAGES=2 timeout 120 python tools/time_layer.py $L 10 >> $O/t.txt 2>&1
cat $O/t.txt
