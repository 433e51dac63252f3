#!/bin/bash
TAG=${1:-bd}; O=gpurun_out/$TAG; mkdir -p $O
timeout 300 python tools/kernel_breakdown.py dsic 8 > $O/dsic.txt 2>&1
timeout 300 python tools/kernel_breakdown.py en 16 > $O/en.txt 2>&1
timeout 300 python tools/kernel_breakdown.py hesic 16 > $O/hesic.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-iters 3 > $O/bench.json 2> $O/bench.err
head -1 $O/dsic.txt; grep -E "hesic::|void " $O/dsic.txt | head -14; head -1 $O/en.txt; grep -E "hesic::|void " $O/en.txt | head -6; head -1 $O/hesic.txt; grep -E "hesic::|void " $O/hesic.txt | head -16; cut -c1-200 $O/bench.json
