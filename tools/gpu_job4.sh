#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"warp_kernel|gaussian_kernel|conv_small_kernel|rowpad_kernel|conv_head_kernel" -c 12 -f -o gpurun_out/ncu_elem python tools/layer_times.py 16 hesic 1 > gpurun_out/ncu_elem.log 2>&1
tail -3 gpurun_out/ncu_elem.log
