#!/bin/bash
mkdir -p gpurun_out/j7
O=gpurun_out/j7
timeout 300 python tools/kernel_breakdown.py en 16 > $O/en.txt 2>&1
timeout 300 python tools/kernel_breakdown.py dsic 8 > $O/dsic.txt 2>&1
head -3 $O/en.txt; head -3 $O/dsic.txt
