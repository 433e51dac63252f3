#!/bin/bash
# first-analysis-layer kernel: correctness against the SIMT path, then timings
TAG=${1:-first}; O=gpurun_out/$TAG; mkdir -p $O
L="3 128 5 2 0 512 512 16 1 0"
timeout 300 python tools/tc_check.py row >> $O/check.txt 2>&1
timeout 300 python tools/tc_check.py edge >> $O/check.txt 2>&1
timeout 120 python tools/time_layer.py $L 10 >> $O/t.txt 2>&1
HESIC_TC_NO_FIRST=1 timeout 120 python tools/time_layer.py $L 10 >> $O/t.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:"conv_tc_first" -s 2 -c 1 -o $O/first python tools/run_layer.py $L 3 > $O/ncu.log 2>&1
cat $O/check.txt | tail -12; cat $O/t.txt
