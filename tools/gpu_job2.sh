#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -f"
# Cin Cout k stride transposed H W B gdn act
timeout 300 $NCU -o gpurun_out/ncu_conv1 python tools/run_layer.py 3 128 5 2 0 512 512 16 1 0 3 > gpurun_out/ncu_conv1.log 2>&1
timeout 300 $NCU -o gpurun_out/ncu_deconv3 python tools/run_layer.py 128 128 5 2 1 128 128 16 2 0 3 > gpurun_out/ncu_deconv3.log 2>&1
timeout 300 $NCU -o gpurun_out/ncu_rgbhead python tools/run_layer.py 128 3 5 2 1 256 256 16 0 0 3 > gpurun_out/ncu_rgbhead.log 2>&1
ls -la gpurun_out
