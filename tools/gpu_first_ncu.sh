#!/bin/bash
TAG=${1:-first_ncu}; O=gpurun_out/$TAG; mkdir -p $O
L="3 128 5 2 0 512 512 16 1 0"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:"conv_tc_first" -s 2 -c 1 -o $O/first python tools/run_layer.py $L 3 > $O/ncu.log 2>&1
tail -3 $O/ncu.log
