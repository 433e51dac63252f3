#!/bin/bash
mkdir -p gpurun_out/j15
O=gpurun_out/j15
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
timeout 300 python tools/dsic_time.py 8 512 512 3 > $O/dsic_time.txt 2>&1; tail -1 $O/dsic_time.txt
timeout 300 python tools/layer_times.py 16 hesic 3 2>&1 | head -1
