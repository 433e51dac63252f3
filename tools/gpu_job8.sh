#!/bin/bash
mkdir -p gpurun_out/j8
O=gpurun_out/j8
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "enhancement" > $O/pytest_en_ops.log 2>&1; echo "rc=$?" >> $O/pytest_en_ops.log
timeout 600 python -m pytest tests/test_gpu_forward.py -q -k "independent_en" > $O/pytest_en_fwd.log 2>&1; echo "rc=$?" >> $O/pytest_en_fwd.log
timeout 300 python tools/kernel_breakdown.py en 16 > $O/en.txt 2>&1
tail -15 $O/pytest_en_ops.log; tail -15 $O/pytest_en_fwd.log; head -1 $O/en.txt
