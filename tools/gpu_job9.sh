#!/bin/bash
mkdir -p gpurun_out/j9
O=gpurun_out/j9
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "enhancement" > $O/pytest_en_ops.log 2>&1; echo "rc=$?" >> $O/pytest_en_ops.log
timeout 600 python -m pytest tests/test_gpu_forward.py -q -k "independent_en" > $O/pytest_en_fwd.log 2>&1; echo "rc=$?" >> $O/pytest_en_fwd.log
timeout 300 python tools/run_en_layer.py 16 5 all > $O/en_layer.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:en_conv_kernel -s 1 -c 1 -o $O/ncu_en_res1 python tools/run_en_layer.py 16 1 res1 > $O/ncu_en_res1.log 2>&1
timeout 300 python tools/kernel_breakdown.py en 16 > $O/en.txt 2>&1
tail -3 $O/pytest_en_ops.log; tail -3 $O/pytest_en_fwd.log; cat $O/en_layer.txt; head -1 $O/en.txt
