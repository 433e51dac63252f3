#!/bin/bash
# ncu --set full (+source) of single layers:  tools/gpu_ncu_layers.sh <tag>
TAG=${1:-ncu}; O=gpurun_out/$TAG; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on -f"
# first analysis layer 3->128 k5 s2 + GDN (16 x 512^2 -> 256^2)
timeout 300 $NCU -k regex:"conv_tc(_pair)?_kernel|conv1_kernel" -s 2 -c 1 -o $O/conv1 python tools/run_layer.py 3 128 5 2 0 512 512 16 1 0 3 > $O/conv1.log 2>&1
# GDN deconv 128->128 k5 s2 + IGDN (16 x 128^2 -> 256^2)
timeout 300 $NCU -k regex:"conv_tc(_pair)?_kernel" -s 2 -c 1 -o $O/deconv3 python tools/run_layer.py 128 128 5 2 1 128 128 16 2 0 3 > $O/deconv3.log 2>&1
for L in "3 128 5 2 0 512 512 16 1 0" "128 128 5 2 1 128 128 16 2 0" "128 128 5 2 0 256 256 16 1 0" "128 960 5 1 0 32 32 16 0 1"; do timeout 120 python tools/time_layer.py $L >> $O/time_layer.txt 2>&1; done
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x > $O/pytest_fullsize.log 2>&1; tail -3 $O/pytest_fullsize.log
tail -2 $O/conv1.log $O/deconv3.log; cat $O/time_layer.txt | tail -20
