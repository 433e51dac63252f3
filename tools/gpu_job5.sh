#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"warp_kernel|conv_small_kernel|gaussian_tile" -c 4 -f -o gpurun_out/ncu_elem2 python tools/layer_times.py 16 hesic 1 > gpurun_out/ncu_elem2.log 2>&1
tail -2 gpurun_out/ncu_elem2.log
