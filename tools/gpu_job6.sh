#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -f"
timeout 300 $NCU -o gpurun_out/ncu_conv1b python tools/run_layer.py 3 128 5 2 0 512 512 16 1 0 3 > gpurun_out/ncu_conv1b.log 2>&1
tail -2 gpurun_out/ncu_conv1b.log
